#!/bin/bash
for lib in default tools/ab/*.so; do
  echo "== $lib"
  if [ $lib != default ]; then export RIP_LIB_PATH=$PWD/$lib; else unset RIP_LIB_PATH; fi
  RIP_FUSED_NPX=4 python tools/prof_fused.py --frames 32 --launches 8
done
