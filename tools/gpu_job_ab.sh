#!/bin/bash
# time every tools/ab/*.so against the default build on iid content: gpu_job_ab.sh TAG [prof_fused args]
TAG=$1; shift
mkdir -p gpurun_out
{
for lib in default tools/ab/*.so; do
  echo "== $lib"
  if [ $lib != default ]; then export RIP_LIB_PATH=$PWD/$lib; else unset RIP_LIB_PATH; fi
  python tools/prof_fused.py --frames 32 --launches 8 "$@"
done
} > gpurun_out/${TAG}_ab.txt 2>&1
cat gpurun_out/${TAG}_ab.txt
