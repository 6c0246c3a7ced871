#!/bin/bash
timeout 600 python -m pytest tests -m gpu -x -q -k "blur or gauss" 2>&1 | tail -n 4
for c in noise alpha255 sky; do python tools/prof_blur.py 17 6.0 16 6 $c; done
for c in noise alpha255; do python tools/prof_blur.py 9 2.5 16 6 $c; done
python tools/prof_blur_artemis.py
python tools/prof_blur_stats.py | grep 17x17
