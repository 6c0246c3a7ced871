#!/bin/bash
timeout 600 python -m pytest tests -m gpu -x -q -k "blur or gauss" 2>&1 | tail -n 3
python tools/prof_blur_small.py
python tools/bench_configs.py 2>&1 | grep -i "gauss"
