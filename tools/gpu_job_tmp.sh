#!/bin/bash
timeout 900 python -m pytest tests -m gpu -x -q -k "fused or sobel or gray" 2>&1 | tail -n 4
for rep in 1 2; do
for lib in default tools/ab/base.so; do
  echo "== $lib"
  if [ $lib != default ]; then export RIP_LIB_PATH=$PWD/$lib; else unset RIP_LIB_PATH; fi
  python tools/prof_fused.py --frames 32 --launches 8
  python tools/prof_fused.py --frames 32 --launches 6 --kind artemis
  python tools/prof_fused.py --frames 32 --launches 6 --kind tulips
  python tools/prof_fused.py --frames 32 --launches 6 --kind flat
  python tools/prof_fused.py --frames 32 --launches 6 --kind halfflat
done
done
