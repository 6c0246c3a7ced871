#!/bin/bash
timeout 600 python -m pytest tests -m gpu -x -q -k "blur or gauss" 2>&1 | tail -n 3
for rep in 1 2; do
for lib in default tools/ab/base.so; do
  echo "== $lib"
  if [ $lib != default ]; then export RIP_LIB_PATH=$PWD/$lib; else unset RIP_LIB_PATH; fi
  for c in noise alpha255 sky; do python tools/prof_blur.py 17 6.0 16 6 $c; done
  for c in noise alpha255; do python tools/prof_blur.py 9 2.5 16 6 $c; done
done
done
unset RIP_LIB_PATH
python tools/prof_blur_artemis.py | grep 17x17
