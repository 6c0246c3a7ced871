#!/bin/bash
# round 2, GPU job W (1 GPU): the build with the reworked stand-alone Gaussian kernels -- full suite, ncu captures of the blur kernels,
# per-config kernel times, bench.py
mkdir -p gpurun_out
O=gpurun_out
( time python -m pytest tests -m gpu -x -q ) > $O/r2w_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/r2w_pytest_gpu.log
tail -n 6 $O/r2w_pytest_gpu.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:blur_stream5 -s 1 -c 1 -o $O/r2w_blur5 -f python tools/prof_blur.py 5 1.0 16 3 alpha255 > $O/r2w_ncu_blur5.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:blur_streamk -s 1 -c 1 -o $O/r2w_blur17 -f python tools/prof_blur.py 17 6.0 16 3 alpha255 > $O/r2w_ncu_blur17.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:blur_streamk -s 1 -c 1 -o $O/r2w_blur9 -f python tools/prof_blur.py 9 2.5 16 3 alpha255 > $O/r2w_ncu_blur9.log 2>&1
python tools/bench_configs.py > $O/r2w_configs.txt 2>&1; cat $O/r2w_configs.txt
python bench.py > $O/r2w_bench.json 2> $O/r2w_bench.err; tail -c 1500 $O/r2w_bench.json
