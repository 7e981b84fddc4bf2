#!/bin/bash
# round 2, call 9 (2 GPUs): multi-GPU bit-identity tests, bench N=2 with the halo_parity leg
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name,clocks.sm --format=csv > gpurun_out/r2c9_smi.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_multi.py -x -q -m gpu > gpurun_out/r2c9_multi.log 2>&1; echo "rc=$?" >> gpurun_out/r2c9_multi.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2c9_bench_n2.json 2> gpurun_out/r2c9_bench_n2.err; echo "bench rc=$?" >> gpurun_out/r2c9_bench_n2.err
tail -5 gpurun_out/r2c9_multi.log; tail -3 gpurun_out/r2c9_bench_n2.err; cat gpurun_out/r2c9_bench_n2.json
