#!/bin/bash
# round 2, call 37: cold path of the fp32 A-A odd kernels as a call (default now) against in line (inl), kernel by kernel; fp64 SRT in the compact form again
mkdir -p gpurun_out
{
for v in cum_f32 srt_f32 srt_f32_inl kbcn4_f32 kbcn4_f32_inl clbm_f32 clbm_f32_inl mrt_f32 mrt_f32_inl; do timeout 300 ./tools/bin/kb_$v 384 20 0 1 0 | grep -v "TMA" | tail -1; done
for v in q9_f32 q9_f32_inl q9clbm_f32 q9clbm_f32_inl; do timeout 300 ./tools/bin/kb_$v 8192 20 0 0 0 | grep -v "TMA" | tail -1; done
timeout 300 ./tools/bin/kb_srt_f64 384 20 0 1 1 | grep -v "TMA" | tail -2
timeout 300 ./tools/bin/kb_q9_f64 8192 20 0 0 1 | grep -v "TMA" | tail -2
} > gpurun_out/r2c37_kbench_f32_odd_per_kernel.txt 2>&1
cat gpurun_out/r2c37_kbench_f32_odd_per_kernel.txt
