#!/bin/bash
# round 2, final 1-GPU job: full suite, smoke, bench.py (default), per-config kernel times, real-image timings, ncu of the streaming KxK kernel
mkdir -p gpurun_out
O=gpurun_out
( time python -m pytest tests -m gpu -x -q ) > $O/r2zq_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/r2zq_pytest_gpu.log
tail -n 6 $O/r2zq_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -n 2
python bench.py > $O/r2zq_bench.json 2> $O/r2zq_bench.err; echo "bench rc=$?"
python bench.py --impl reference --steps 3 --warmup 1 > $O/r2zq_bench_reference.json 2> $O/r2zq_bench_reference.err; echo "bench reference rc=$?"; tail -c 400 $O/r2zq_bench_reference.json
python tools/bench_configs.py > $O/r2zq_configs.txt 2>&1; cat $O/r2zq_configs.txt
python tools/prof_blur_artemis.py > $O/r2zq_blur_real_images.txt 2>&1; python tools/prof_blur_stats.py >> $O/r2zq_blur_real_images.txt 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:blur_streamk -s 1 -c 1 -o $O/r2zq_blur17 -f python tools/prof_blur.py 17 6.0 16 3 alpha255 > $O/r2zq_ncu_blur17.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:blur_stream5 -s 1 -c 1 -o $O/r2zq_blur5 -f python tools/prof_blur.py 5 1.0 16 3 alpha255 > $O/r2zq_ncu_blur5.log 2>&1
{
for size in 1920x1080 3840x2160; do for m in FUSED GAUSSIAN; do tools/rip_headless.bin stream $size --frames 200 --inflight 3 --method $m; done; done
} > $O/r2zq_headless_stream.txt 2>&1; cat $O/r2zq_headless_stream.txt
