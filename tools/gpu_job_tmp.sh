#!/bin/bash
python -m pytest tests -m gpu -x -q 2>&1 | tail -n 5
for kind in uniform flat zero; do python tools/prof_fused.py --frames 32 --kind $kind --launches 6; done
