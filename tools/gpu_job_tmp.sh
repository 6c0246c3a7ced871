#!/bin/bash
for rep in 1 2; do
for seg in 0 128 240; do
for lib in default tools/ab/minb5c.so; do
  echo "== $lib seg $seg"
  if [ $lib != default ]; then export RIP_LIB_PATH=$PWD/$lib; else unset RIP_LIB_PATH; fi
  if [ $seg != 0 ]; then export RIP_FUSED_SEG=$seg; else unset RIP_FUSED_SEG; fi
  python tools/prof_fused.py --frames 32 --launches 8
done
done
done
