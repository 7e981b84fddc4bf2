#!/bin/bash
# round 2, call 43 (last): the final library through the full GPU suite; SRT_MODIF_FORCE rows of the engine table
mkdir -p gpurun_out
timeout 280 python -m pytest tests -q -m gpu > gpurun_out/r2c43_gpu_suite.log 2>&1; echo "rc=$?" >> gpurun_out/r2c43_gpu_suite.log
timeout 60 python tools/bench_matrix.py --operator SRT_MODIF --out gpurun_out/r2c43_bench_matrix_srtmf.md > gpurun_out/r2c43_bench_matrix_srtmf.log 2>&1
tail -3 gpurun_out/r2c43_gpu_suite.log; cat gpurun_out/r2c43_bench_matrix_srtmf.log
