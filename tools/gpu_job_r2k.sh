#!/bin/bash
# round 2, 8-GPU job: in-process multi-device tests (one rip_ctx over all 8 devices), PCIe probe at 1/2/4/8, bench.py at N=8 and N=2
mkdir -p gpurun_out
O=gpurun_out
nvidia-smi -L > $O/r2k_gpus.txt 2>&1
nvidia-smi topo -m >> $O/r2k_gpus.txt 2>&1
( time python -m pytest tests/test_multi_device.py tests/test_host_pipeline.py -m gpu -x -q -s ) > $O/r2k_pytest_multi_device.log 2>&1; echo "pytest rc=$?" >> $O/r2k_pytest_multi_device.log
tail -n 6 $O/r2k_pytest_multi_device.log
timeout 300 tools/pcie_probe.bin 0.3 > $O/r2k_pcie_probe_8gpu.txt 2>&1
cat $O/r2k_pcie_probe_8gpu.txt
for N in 8 2; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 > $O/r2k_bench_${N}gpu.json 2> $O/r2k_bench_${N}gpu.err
echo "bench N=$N rc=$?"; tail -c 2500 $O/r2k_bench_${N}gpu.json; tail -n 3 $O/r2k_bench_${N}gpu.err
done
