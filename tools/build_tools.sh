#!/bin/bash
# Build the stand-alone probes under tools/ for sm_100a (binaries are git-ignored, they travel with gpurun).
set -e
cd "$(dirname "$0")"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
for t in pcie_probe pipe_probe tma_probe; do
    [ -f $t.cu ] || continue
    if [ ! -f $t.bin ] || [ $t.cu -nt $t.bin ]; then
        $NVCC -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -ccbin /usr/bin/g++ -Xcompiler -pthread -o $t.bin $t.cu -lcuda 2>&1 | grep -v "^$" || true
    fi
done
ls -la *.bin
