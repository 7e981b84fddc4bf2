#!/bin/bash
# round 2, call 39: the fp32 KBC A-B kernels at 5 (default) and 4 CTAs per SM, model by model (KBC_N1 fell to 58 % in the final table while N4 sits at 71 %)
mkdir -p gpurun_out
{
for v in kbcn1_f32 kbcn1_f32_ab4 kbcn2_f32 kbcn2_f32_ab4 kbcc1_f32 kbcc1_f32_ab4 kbcn4_f32_ab4 kbcc4_f32_ab4; do timeout 300 ./tools/bin/kb_$v 384 20 0 1 0 | grep -v "TMA" | tail -1; done
} > gpurun_out/r2c39_kbench_kbc_f32_ab.txt 2>&1
cat gpurun_out/r2c39_kbench_kbc_f32_ab.txt
