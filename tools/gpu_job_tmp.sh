#!/bin/bash
timeout 600 python -m pytest tests -m gpu -x -q -k "blur or gauss" 2>&1 | tail -n 3
python tools/prof_blur_artemis.py
python tools/prof_blur_small.py | grep "17x17\|9x9"
