#!/bin/bash
timeout 600 python -m pytest tests -m gpu -x -q -k "blur or gauss" 2>&1 | tail -n 5
for c in noise alpha255 sky; do timeout 120 python tools/prof_blur.py 17 6.0 16 6 $c; done
for c in noise alpha255; do timeout 120 python tools/prof_blur.py 9 2.5 16 6 $c; done
for c in noise alpha255; do timeout 120 python tools/prof_blur.py 17 6.0 1 6 $c; done
