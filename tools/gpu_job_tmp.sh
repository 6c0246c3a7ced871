#!/bin/bash
for rep in 1 2; do
for lib in default tools/ab/nb7.so; do
  echo "== $lib"
  if [ $lib != default ]; then export RIP_LIB_PATH=$PWD/$lib; else unset RIP_LIB_PATH; fi
  python tools/prof_fused.py --op sobel --frames 64 --w 1920 --h 1080 --launches 10
  python tools/prof_fused.py --op sobel --frames 32 --launches 10
done
done
