#!/bin/bash
# round 2, call 29: the cumulant operator built with the 2017 / antialias switches in its default-arithmetic reorganisation
# (collide_cum_switches) against the reference statement order; KBC with the final register caps; the whole kernel-family table through
# the engine; the GPU parity tests
mkdir -p gpurun_out
{
for v in cum2017aa_f64_ref cum2017aa_f64 cum2017aa_f32_ref cum2017aa_f32 cum2017_f64 cum2017_f32 cumaa_f32 kbcn4_f64 kbcc4_f64 kbcn4_f32 kbcc4_f32 clbm_f64 clbm_f32; do timeout 300 ./tools/bin/kb_$v 384 20 0 1 0 | grep -v "TMA" | tail -1; done
} > gpurun_out/r2c29_kbench_cum_switches_kbc.txt 2>&1
timeout 1200 python -m pytest tests/test_gpu_parity.py -q -m gpu -x > gpurun_out/r2c29_parity.log 2>&1; echo "rc=$?" >> gpurun_out/r2c29_parity.log
timeout 1500 python tools/bench_matrix.py --out gpurun_out/r2c29_bench_matrix.md > gpurun_out/r2c29_bench_matrix.log 2>&1; echo "rc=$?" >> gpurun_out/r2c29_bench_matrix.log
cat gpurun_out/r2c29_kbench_cum_switches_kbc.txt; tail -4 gpurun_out/r2c29_parity.log; tail -30 gpurun_out/r2c29_bench_matrix.log
