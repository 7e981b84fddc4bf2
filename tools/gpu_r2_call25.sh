#!/bin/bash
# round 2, call 25 (8 GPUs): bench N=8 with the halo-parity and channel legs, 3/4/8-slab bit-identity tests, box topology
mkdir -p gpurun_out
{ nvidia-smi topo -m; echo; lscpu | grep -E "Model name|Socket|NUMA|^CPU\(s\)"; echo; (numactl -H 2>/dev/null || cat /sys/devices/system/node/node*/meminfo 2>/dev/null | grep MemTotal); echo; cat /proc/self/status | grep -i cpus_allowed_list; nvidia-smi --query-gpu=index,pci.bus_id,pcie.link.gen.current,pcie.link.width.current --format=csv; } > gpurun_out/r2c25_topology.txt 2>&1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r2c25_bench_n8.json 2> gpurun_out/r2c25_bench_n8.err; echo "bench rc=$?" >> gpurun_out/r2c25_bench_n8.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 4 --steps 20 --warmup 5 > gpurun_out/r2c25_bench_n4.json 2> gpurun_out/r2c25_bench_n4.err; echo "bench rc=$?" >> gpurun_out/r2c25_bench_n4.err
timeout 900 python -m pytest tests/test_gpu_multi.py -q -m gpu -k "n_gpus or late_rank" > gpurun_out/r2c25_multi.log 2>&1; echo "rc=$?" >> gpurun_out/r2c25_multi.log
timeout 300 python -m pytest tests/test_dropin_solvers.py -q -m gpu -k "two_ranks" > gpurun_out/r2c25_dropin2.log 2>&1; echo "rc=$?" >> gpurun_out/r2c25_dropin2.log
cat gpurun_out/r2c25_topology.txt; tail -4 gpurun_out/r2c25_bench_n8.err; cat gpurun_out/r2c25_bench_n8.json; tail -3 gpurun_out/r2c25_multi.log gpurun_out/r2c25_dropin2.log; tail -2 gpurun_out/r2c25_bench_n4.err
