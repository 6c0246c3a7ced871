#!/bin/bash
python -m pytest tests -m gpu -x -q 2>&1 | tail -n 25
for c in noise alpha255 sky; do python tools/prof_blur.py 5 1.0 16 4 $c; done
