#!/bin/bash
# round 2, call 41 (2 GPUs): the final library through the multi-GPU bit-identity tests and the N = 2 bench line (halo_parity, channel_strong)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_multi.py -x -q -m gpu > gpurun_out/r2c41_multi.log 2>&1; echo "rc=$?" >> gpurun_out/r2c41_multi.log
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 20 --warmup 5 --no-extras --no-cpu-baseline > gpurun_out/r2c41_bench_n2.json 2> gpurun_out/r2c41_bench_n2.err; echo "bench rc=$?" >> gpurun_out/r2c41_bench_n2.err
tail -3 gpurun_out/r2c41_multi.log; tail -2 gpurun_out/r2c41_bench_n2.err | cut -c1-300; cat gpurun_out/r2c41_bench_n2.json | cut -c1-600
