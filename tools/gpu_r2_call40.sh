#!/bin/bash
# round 2, call 40: library with the per-model fp32 KBC A-B register caps -- GPU parity tests and the KBC / cumulant-switch rows of the engine table
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_physics_and_edges.py -q -m gpu > gpurun_out/r2c40_parity.log 2>&1; echo "rc=$?" >> gpurun_out/r2c40_parity.log
timeout 900 python tools/bench_matrix.py --only ext2 --out gpurun_out/r2c40_bench_matrix_ext2.md > gpurun_out/r2c40_bench_matrix_ext2.log 2>&1; echo "rc=$?" >> gpurun_out/r2c40_bench_matrix_ext2.log
tail -3 gpurun_out/r2c40_parity.log; cat gpurun_out/r2c40_bench_matrix_ext2.log
