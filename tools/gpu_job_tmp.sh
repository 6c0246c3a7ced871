#!/bin/bash
python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -n 3
for kind in uniform letterbox halfflat flat zero; do python tools/prof_fused.py --frames 32 --kind $kind --launches 6; done
bash tools/gpu_job_ab.sh r2i
