#!/bin/bash
# GPU-side collection of the round's profiling evidence (run under gpurun; artefacts land in gpurun_out/).
#   usage: bash tools/run_profiles.sh r02
TAG=${1:-r02}
O=gpurun_out
mkdir -p $O
NCU="ncu --clock-control none"
# 1. launch list of the bench command (cold-cache, serialised: compare SHARES)
$NCU --metrics gpu__time_duration.sum -s 60 -c 400 --csv --log-file $O/${TAG}_launches.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extra-legs > $O/${TAG}_launches_bench.log 2>&1
# 2. full capture of every kernel of the 16-bit path (one half-batch of 128 images per launch)
$NCU --set full --import-source on -s 18 -c 9 -o $O/${TAG}_full -f \
    python tools/profile_once.py --batch 256 --iters 2 > $O/${TAG}_full.log 2>&1
# 3. fp32 path kernels (batch 8)
$NCU --set full --import-source on -c 30 -o $O/${TAG}_fp32 -f \
    python tools/profile_once.py --batch 8 --iters 1 --precision fp32 > $O/${TAG}_fp32.log 2>&1
# 3b. fp32-class tensor-core path (split-fp16 layers), one half-batch of 128 images per launch
$NCU --set full --import-source on -s 20 -c 10 -o $O/${TAG}_fp32tc -f \
    python tools/profile_once.py --batch 256 --iters 2 --precision fp32tc > $O/${TAG}_fp32tc.log 2>&1
# 4. front-end kernels, 300 / 600 variants
$NCU --set full --import-source on -k regex:"crop_resize|yuv420|jpeg_|huff_|prep_u8|chunked_to_f32|avgpool|join_kernel|dense_tail|conv3x3" -c 56 \
    -o $O/${TAG}_front -f python tools/profile_front.py > $O/${TAG}_front.log 2>&1
tail -n 2 $O/${TAG}_full.log $O/${TAG}_fp32.log $O/${TAG}_front.log
# gpurun brings back at most 64 MiB: keep the raw metric pages as CSV (what tools/summarize_profiles.py reads) and, of the
# reports themselves, only the fused block's (source-level analysis happens off the box)
for s in full fp32 fp32tc front; do
  ncu -i $O/${TAG}_$s.ncu-rep --page raw --csv > $O/${TAG}_$s.raw.csv 2> /dev/null
done
ncu -i $O/${TAG}_full.ncu-rep --page source --csv --print-source sass -k regex:block2 > $O/${TAG}_block2_source.csv 2> /dev/null
rm -f $O/${TAG}_full.ncu-rep $O/${TAG}_fp32.ncu-rep $O/${TAG}_fp32tc.ncu-rep $O/${TAG}_front.ncu-rep
du -sh $O
