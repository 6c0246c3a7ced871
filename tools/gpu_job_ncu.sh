#!/bin/bash
# one ncu --set full capture of the fused kernel on 32 4K RGB frames: gpu_job_ncu.sh TAG [extra prof_fused args]
TAG=$1; shift
mkdir -p gpurun_out
python tools/prof_fused.py --frames 32 --launches 4 "$@" > gpurun_out/${TAG}_plain.log 2>&1 || exit 1
cat gpurun_out/${TAG}_plain.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fused_x2 -s 1 -c 1 -o gpurun_out/${TAG} -f python tools/prof_fused.py --frames 32 --launches 3 "$@" > gpurun_out/${TAG}_ncu.log 2>&1
tail -n 3 gpurun_out/${TAG}_ncu.log
