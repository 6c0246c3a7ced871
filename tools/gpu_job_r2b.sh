#!/bin/bash
# round 2, GPU job B: the ring-read fused kernel (x3): parity, then timings of three register budgets on every content kind
mkdir -p gpurun_out
O=gpurun_out
( time python -m pytest tests -m gpu -x -q ) > $O/r2b_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/r2b_pytest_gpu.log
tail -n 15 $O/r2b_pytest_gpu.log
{
for lib in default; do
  echo "== $lib"
  if [ $lib != default ]; then export RIP_LIB_PATH=$PWD/$lib; else unset RIP_LIB_PATH; fi
  for kind in uniform smooth letterbox halfflat flat zero; do python tools/prof_fused.py --frames 32 --kind $kind --launches 6; done
  python tools/prof_fused.py --frames 32 --fmt rgba --launches 6
  python tools/prof_fused.py --frames 32 --fmt gray --launches 6
  python tools/prof_fused.py --frames 32 --op sobel --launches 6
done
unset RIP_LIB_PATH
} > $O/r2b_timings.txt 2>&1
cat $O/r2b_timings.txt
