#!/bin/bash
# Round-end verification on one B200 (run under gpurun): GPU test suite, smoke, bench lines, JPEG bench, launch list
# of the JPEG kernels and the `ncu --set full` capture of the front-end kernels.  Artefacts land in gpurun_out/.
TAG=${1:-r02}
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee $O/${TAG}_gpu_tests.log
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -6 | tee $O/${TAG}_smoke.log
timeout 600 python bench.py > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err; tail -c 400 $O/${TAG}_bench.json
timeout 300 python tools/bench_jpeg.py 384 > $O/${TAG}_jpeg_bench.json 2> $O/${TAG}_jpeg_bench.err; cut -c1-900 $O/${TAG}_jpeg_bench.json
ncu --clock-control none --metrics gpu__time_duration.sum -k regex:"huff_|jpeg_|crop_resize" --csv \
    --log-file $O/${TAG}_jpeg_launches.csv python tools/experiments/jpeg_kernels.py 2>&1 | tail -1
ncu --clock-control none --set full --import-source on \
    -k regex:"crop_resize|yuv420|jpeg_|huff_|prep_u8|chunked_to_f32|avgpool|join_kernel|dense_tail|conv3x3" -c 56 \
    -o $O/${TAG}_front -f python tools/profile_front.py > $O/${TAG}_front.log 2>&1
ncu -i $O/${TAG}_front.ncu-rep --page raw --csv > $O/${TAG}_front.raw.csv 2> /dev/null
rm -f $O/${TAG}_front.ncu-rep
tail -n 2 $O/${TAG}_front.log
