#!/bin/bash
# round 2, GPU job A (1 GPU): full GPU test suite on the new host pipeline, PCIe probe, sanitizer logs, blur profiles,
# RGBA / content timings of the fused kernel as it stood at the start of the round.
mkdir -p gpurun_out
O=gpurun_out
( time python -m pytest tests -m gpu -x -q ) > $O/r2a_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/r2a_pytest_gpu.log
tools/pcie_probe.bin 0.3 > $O/r2a_pcie_probe_1gpu.txt 2>&1
{
for kind in uniform smooth letterbox halfflat flat zero; do python tools/prof_fused.py --frames 32 --kind $kind --launches 6; done
python tools/prof_fused.py --frames 32 --fmt rgba --launches 6
python tools/prof_fused.py --frames 32 --fmt rgba --op sobel --launches 6
python tools/prof_blur.py 5 1.0 16 6
python tools/prof_blur.py 17 6.0 16 4
python tools/prof_blur.py 17 6.0 1 4
} > $O/r2a_timings.txt 2>&1
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python tools/sanitize_run.py > $O/r2a_sanitizer_memcheck.log 2>&1; echo "rc=$?" >> $O/r2a_sanitizer_memcheck.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python tools/sanitize_run.py > $O/r2a_sanitizer_racecheck.log 2>&1; echo "rc=$?" >> $O/r2a_sanitizer_racecheck.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:blur -c 2 -o $O/r2a_blur5 -f python tools/prof_blur.py 5 1.0 16 2 > $O/r2a_ncu_blur5.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:blur -c 2 -o $O/r2a_blur17 -f python tools/prof_blur.py 17 6.0 16 2 > $O/r2a_ncu_blur17.log 2>&1
tail -3 $O/r2a_pytest_gpu.log; cat $O/r2a_timings.txt; tail -2 $O/r2a_sanitizer_memcheck.log $O/r2a_sanitizer_racecheck.log
