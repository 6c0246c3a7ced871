#!/bin/bash
# round 2, GPU job E: segment-length sweep of the fused kernel (wave quantisation) and A/B builds
mkdir -p gpurun_out
O=gpurun_out
{
for seg in 90 98 108 120 128 135 144 160 180 216 240 270; do
  echo "== seg $seg"; RIP_FUSED_SEG=$seg python tools/prof_fused.py --frames 32 --launches 8
done
} > $O/r2e_seg.txt 2>&1
cat $O/r2e_seg.txt
bash tools/gpu_job_ab.sh r2e
for lib in minb5 acc; do echo "== $lib seg 120/240"; for seg in 120 240; do RIP_LIB_PATH=$PWD/tools/ab/$lib.so RIP_FUSED_SEG=$seg python tools/prof_fused.py --frames 32 --launches 8; done; done 2>&1 | tee $O/r2e_ab_seg.txt
