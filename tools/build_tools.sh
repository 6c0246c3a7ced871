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
# the headless driver of the host classes (g++ against the in-tree libraries)
PKG=../opencl-development-real-time-image-processing_b200
if [ ! -f rip_headless.bin ] || [ rip_headless.cpp -nt rip_headless.bin ] || [ $PKG/librip_host.so -nt rip_headless.bin ]; then
    /usr/bin/g++ -O2 -std=c++17 -ffp-contract=off -I $PKG/host -I ../include -I /usr/local/cuda/include rip_headless.cpp -o rip_headless.bin \
        -L $PKG -lrip_host -lrip_cuda -Wl,-rpath,'$ORIGIN/'$PKG -pthread
fi
ls -la *.bin
