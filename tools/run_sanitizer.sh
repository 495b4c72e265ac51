#!/bin/bash
# compute-sanitizer runs kept as evidence (gpurun_out/<tag>_sanitizer_*.log): memcheck over the shapes the over-read
# slack has to cover (batch 1, 3, 65, 255 at 224; batch 2 at 300 and 600), racecheck + synccheck on a small batch.
TAG=${1:-r02}
O=gpurun_out
CS=/usr/local/cuda/bin/compute-sanitizer
for b in 1 3 65 255; do
  timeout 600 $CS --tool memcheck --error-exitcode 3 python tools/profile_once.py --batch $b --iters 1 \
      > $O/${TAG}_sanitizer_memcheck_b$b.log 2>&1; echo "memcheck batch $b rc=$?" | tee -a $O/${TAG}_sanitizer_summary.log
done
for s in 300 600; do
  timeout 600 $CS --tool memcheck --error-exitcode 3 python tools/profile_once.py --batch 2 --iters 1 --side $s \
      > $O/${TAG}_sanitizer_memcheck_s$s.log 2>&1; echo "memcheck side $s rc=$?" | tee -a $O/${TAG}_sanitizer_summary.log
done
timeout 900 $CS --tool racecheck --error-exitcode 3 python tools/profile_once.py --batch 3 --iters 1 \
    > $O/${TAG}_sanitizer_racecheck_b3.log 2>&1; echo "racecheck batch 3 rc=$?" | tee -a $O/${TAG}_sanitizer_summary.log
timeout 900 $CS --tool synccheck --error-exitcode 3 python tools/profile_once.py --batch 3 --iters 1 \
    > $O/${TAG}_sanitizer_synccheck_b3.log 2>&1; echo "synccheck batch 3 rc=$?" | tee -a $O/${TAG}_sanitizer_summary.log
timeout 600 $CS --tool memcheck --error-exitcode 3 python tools/profile_front.py \
    > $O/${TAG}_sanitizer_memcheck_front.log 2>&1; echo "memcheck front end rc=$?" | tee -a $O/${TAG}_sanitizer_summary.log
grep -H "ERROR SUMMARY" $O/${TAG}_sanitizer_*.log
for f in $O/${TAG}_sanitizer_*.log; do  # keep the logs small enough to travel back
  head -c 200000 $f > $f.tmp && mv $f.tmp $f
done
