#!/bin/bash
# round 2, call 31: why the engine's D2Q9 and D3Q19-SRT fp64 rows of bench_matrix_r2 sit 10-19 % below round 1: the kernels one by one and
# chained (kbench), with the round-2 features switched off one at a time; KBC fp32 with the IEEE division kept in the A-B kernel
mkdir -p gpurun_out
{
for v in q9_f64 q9_f64_call q9_f64_r1hints q9_f64_cpt1 q9_f64; do timeout 300 ./tools/bin/kb_$v 8192 20 0 0 1 | grep -v "TMA" | tail -2; done
for v in q19srt_f64 q19srt_f64_ws0 q19srt_f64_call q19srt_f64; do timeout 300 ./tools/bin/kb_$v 384 20 0 1 1 | grep -v "TMA" | tail -2; done
timeout 300 ./tools/bin/kb_kbcn4_f32 384 20 0 1 0 | grep -v "TMA" | tail -1
} > gpurun_out/r2c31_kbench_q9_q19srt.txt 2>&1
timeout 600 python tools/bench_matrix.py --only D2Q9 --out gpurun_out/r2c31_bench_matrix_d2q9.md > gpurun_out/r2c31_bench_matrix_d2q9.log 2>&1
cat gpurun_out/r2c31_kbench_q9_q19srt.txt; cat gpurun_out/r2c31_bench_matrix_d2q9.log
