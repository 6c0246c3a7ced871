#!/bin/bash
# round 2, GPU job L (1 GPU): full suite, the headless program (per-image table + streaming loop at 1080p / 4K), sanitizer logs
mkdir -p gpurun_out
O=gpurun_out
( time python -m pytest tests -m gpu -x -q ) > $O/r2l_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/r2l_pytest_gpu.log
tail -n 6 $O/r2l_pytest_gpu.log
{
mkdir -p /tmp/rip_imgs
tools/rip_headless.bin images /tmp/rip_imgs --iterations 20 --csv $O/r2l_results_extended.csv --hbm-peak 6550.4 --synthetic 75x75 --synthetic 427x240 --synthetic 640x512 --synthetic 683x1023 --synthetic 1920x1080 --synthetic 3840x2160
for size in 1920x1080 3840x2160; do
  for m in FUSED EDGE GRAYSCALE GAUSSIAN; do tools/rip_headless.bin stream $size --frames 200 --inflight 3 --method $m; done
done
tools/rip_headless.bin stream 1920x1080 --frames 100 --inflight 3 --method GAUSSIAN --ksize 17 --sigma 6
} > $O/r2l_headless.txt 2>&1
cat $O/r2l_headless.txt
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python tools/sanitize_run.py > $O/r2l_sanitizer_memcheck.log 2>&1; echo "rc=$?" >> $O/r2l_sanitizer_memcheck.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python tools/sanitize_run.py > $O/r2l_sanitizer_racecheck.log 2>&1; echo "rc=$?" >> $O/r2l_sanitizer_racecheck.log
tail -n 4 $O/r2l_sanitizer_memcheck.log $O/r2l_sanitizer_racecheck.log
