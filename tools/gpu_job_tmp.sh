#!/bin/bash
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -n 5
for rep in 1 2; do
for lib in default tools/ab/base.so; do
  echo "== $lib"
  if [ $lib != default ]; then export RIP_LIB_PATH=$PWD/$lib; else unset RIP_LIB_PATH; fi
  python tools/prof_fused.py --frames 32 --launches 8
  python tools/prof_fused.py --frames 32 --launches 8 --fmt rgba
  python tools/prof_fused.py --frames 32 --launches 8 --fmt gray
  python tools/prof_fused.py --op sobel --frames 32 --launches 8
  python tools/prof_fused.py --op sobel --frames 64 --w 1920 --h 1080 --launches 8
  python tools/prof_fused.py --frames 32 --launches 6 --kind flat
done
done
