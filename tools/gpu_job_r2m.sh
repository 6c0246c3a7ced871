#!/bin/bash
# round 2, GPU job M (1 GPU): the shipped build -- full suite, ncu captures (fused with source, blur kernels), launch list of bench.py,
# per-config kernel times, streaming loop through the host classes
mkdir -p gpurun_out
O=gpurun_out
( time python -m pytest tests -m gpu -x -q ) > $O/r2m_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/r2m_pytest_gpu.log
tail -n 6 $O/r2m_pytest_gpu.log
bash tools/gpu_job_ncu.sh r2m_fused
python bench.py --steps 2 --warmup 3 --no-extras --no-cpu-baseline > $O/r2m_bench_plain.json 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r2m_bench_launches.csv python bench.py --steps 2 --warmup 3 --no-extras --no-cpu-baseline > $O/r2m_bench_ncu.log 2>&1
python tools/prof_blur.py 5 1.0 16 4 && timeout 300 ncu --set full --clock-control none --import-source on -k regex:blur -c 1 -o $O/r2m_blur5 -f python tools/prof_blur.py 5 1.0 16 2 > $O/r2m_ncu_blur5.log 2>&1
python tools/prof_blur.py 17 6.0 16 4 && timeout 300 ncu --set full --clock-control none --import-source on -k regex:blur -c 1 -o $O/r2m_blur17 -f python tools/prof_blur.py 17 6.0 16 2 > $O/r2m_ncu_blur17.log 2>&1
python tools/bench_configs.py > $O/r2m_configs.txt 2>&1; cat $O/r2m_configs.txt
{
for size in 1920x1080 3840x2160; do for m in FUSED GAUSSIAN; do tools/rip_headless.bin stream $size --frames 200 --inflight 3 --method $m; done; done
} > $O/r2m_headless_stream.txt 2>&1; cat $O/r2m_headless_stream.txt
