#!/bin/bash
# round 2, 8-GPU job of the shipped build (cp.async row fetch): in-process multi-device tests over one context of 8 devices, bench.py at N = 8, 4, 2
mkdir -p gpurun_out
O=gpurun_out
nvidia-smi -L > $O/r2zi_gpus.txt 2>&1
( time python -m pytest tests/test_multi_device.py tests/test_host_pipeline.py -m gpu -x -q -s ) > $O/r2zi_pytest_multi_device.log 2>&1; echo "pytest rc=$?" >> $O/r2zi_pytest_multi_device.log
tail -n 6 $O/r2zi_pytest_multi_device.log
for N in 8 4 2; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 > $O/r2zi_bench_${N}gpu.json 2> $O/r2zi_bench_${N}gpu.err
echo "bench N=$N rc=$?"; tail -c 1200 $O/r2zi_bench_${N}gpu.json; tail -n 3 $O/r2zi_bench_${N}gpu.err
done
