#!/bin/bash
# round 2, call 42: D2Q9 CLBM in its default-arithmetic form (raw moments once, shared cascade sums, separable source): kernel rates, the D2Q9 rows of
# the engine table, GPU parity and drop-in tests (sim2d_3 runs D2Q9_CLBM)
mkdir -p gpurun_out
{
for v in q9clbm_f64 q9clbm_f32; do timeout 300 ./tools/bin/kb_$v 8192 20 0 0 1 | grep -v "TMA" | tail -2; done
} > gpurun_out/r2c42_kbench_q9clbm.txt 2>&1
timeout 600 python tools/bench_matrix.py --only D2Q9 --out gpurun_out/r2c42_bench_matrix_d2q9.md > gpurun_out/r2c42_bench_matrix_d2q9.log 2>&1
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_dropin_solvers.py -q -m gpu > gpurun_out/r2c42_parity.log 2>&1; echo "rc=$?" >> gpurun_out/r2c42_parity.log
cat gpurun_out/r2c42_kbench_q9clbm.txt gpurun_out/r2c42_bench_matrix_d2q9.log; tail -3 gpurun_out/r2c42_parity.log
