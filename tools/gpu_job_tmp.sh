#!/bin/bash
python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "fused or sobel or selftest" 2>&1 | tail -n 3
for kind in uniform flat halfflat zero; do python tools/prof_fused.py --frames 32 --kind $kind --launches 6; done
python tools/prof_fused.py --frames 32 --fmt rgba --launches 6
python tools/prof_fused.py --frames 32 --fmt gray --launches 6
bash tools/gpu_job_ncu.sh r2g_x3
