#!/bin/bash
# round 2, GPU job ZH (1 GPU): the shipped build after the cp.async row fetch -- full suite, ncu captures (fused with source, 5x5 blur), launch list
# of bench.py, per-config kernel times, bench.py, streaming loop through the host classes
mkdir -p gpurun_out
O=gpurun_out
( time python -m pytest tests -m gpu -x -q ) > $O/r2zh_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/r2zh_pytest_gpu.log
tail -n 6 $O/r2zh_pytest_gpu.log
bash tools/gpu_job_ncu.sh r2zh_fused
python bench.py --steps 2 --warmup 3 --no-extras --no-cpu-baseline > $O/r2zh_bench_plain.json 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r2zh_bench_launches.csv python bench.py --steps 2 --warmup 3 --no-extras --no-cpu-baseline > $O/r2zh_bench_ncu.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:blur_stream5 -s 1 -c 1 -o $O/r2zh_blur5 -f python tools/prof_blur.py 5 1.0 16 3 alpha255 > $O/r2zh_ncu_blur5.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:fused_x2 -s 1 -c 1 -o $O/r2zh_sobel -f python tools/prof_fused.py --op sobel --frames 64 --w 1920 --h 1080 --launches 3 > $O/r2zh_ncu_sobel.log 2>&1
python tools/bench_configs.py > $O/r2zh_configs.txt 2>&1; cat $O/r2zh_configs.txt
python bench.py > $O/r2zh_bench.json 2> $O/r2zh_bench.err; tail -c 600 $O/r2zh_bench.json
{
for size in 1920x1080 3840x2160; do for m in FUSED GAUSSIAN; do tools/rip_headless.bin stream $size --frames 200 --inflight 3 --method $m; done; done
} > $O/r2zh_headless_stream.txt 2>&1; cat $O/r2zh_headless_stream.txt
