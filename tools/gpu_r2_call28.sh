#!/bin/bash
# round 2, call 28: KBC and CLBM in their default-arithmetic reorganisation (collide_kbc_fast, collide_clbm_fast): kbench variants
# (register caps, the reference statement order for comparison), then the GPU parity tests that exercise them
mkdir -p gpurun_out
{
for v in kbcn4_f64_ref kbcn4_f64 kbcn4_f64_odd3 kbcn4_f64_mb4 kbcc4_f64 kbcn1_f64 kbcn4_f32 kbcn4_f32_mb5 kbcn4_f32_mb3 clbm_f64 clbm_f32; do timeout 300 ./tools/bin/kb_$v 384 20 0 1 0 | grep -v "TMA" | tail -1; done
} > gpurun_out/r2c28_kbench_kbc_clbm.txt 2>&1
timeout 1200 python -m pytest tests/test_gpu_parity.py -q -m gpu -x > gpurun_out/r2c28_parity.log 2>&1; echo "rc=$?" >> gpurun_out/r2c28_parity.log
cat gpurun_out/r2c28_kbench_kbc_clbm.txt; tail -4 gpurun_out/r2c28_parity.log
