#!/bin/bash
# round 2, GPU job D (1 GPU): state of the tree after re-entry: full GPU suite, content timings of the x3 kernel, ncu capture
mkdir -p gpurun_out
O=gpurun_out
( time python -m pytest tests -m gpu -x -q ) > $O/r2d_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/r2d_pytest_gpu.log
tail -n 8 $O/r2d_pytest_gpu.log
{
for kind in uniform smooth letterbox halfflat flat zero; do python tools/prof_fused.py --frames 32 --kind $kind --launches 6; done
python tools/prof_fused.py --frames 32 --fmt rgba --launches 6
python tools/prof_fused.py --frames 32 --fmt gray --launches 6
python tools/prof_fused.py --frames 32 --op sobel --launches 6
python tools/prof_blur.py 5 1.0 16 6
python tools/prof_blur.py 17 6.0 16 4
} > $O/r2d_timings.txt 2>&1
cat $O/r2d_timings.txt
bash tools/gpu_job_ncu.sh r2d_x3
python bench.py --steps 10 --warmup 3 > $O/r2d_bench.json 2> $O/r2d_bench.err; tail -c 3000 $O/r2d_bench.json
