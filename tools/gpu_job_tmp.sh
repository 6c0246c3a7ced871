#!/bin/bash
python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "blur" 2>&1 | tail -n 4
python tools/prof_blur.py 5 1.0 16 6
python tools/prof_blur.py 17 6.0 16 4
python tools/prof_blur.py 17 6.0 1 4
python tools/prof_blur.py 9 2.5 16 4
python tools/prof_blur.py 3 0.8 16 4
RIP_BLUR_TILED=1 python tools/prof_blur.py 5 1.0 16 4
