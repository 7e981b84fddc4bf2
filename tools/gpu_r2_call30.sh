#!/bin/bash
# round 2, call 30: KBC fp32 A-B at 4 / 5 CTAs per SM with and without the hardware reciprocal (the 5-CTA build lost 23 % between calls 28
# and 29); KBC fp64 A-A odd with the streaming cache hints of the even kernel
mkdir -p gpurun_out
{
for v in kbcn4_f32_ab4_rcp kbcn4_f32_ab5_rcp kbcn4_f32_ab4_norcp kbcn4_f32_ab5_norcp kbcn4_f32_ab5_rcp kbcn4_f32_ab5_norcp kbcn4_f64 kbcn4_f64_oddcg; do timeout 300 ./tools/bin/kb_$v 384 20 0 1 0 | grep -v "TMA" | tail -1; done
} > gpurun_out/r2c30_kbench_kbc_variants.txt 2>&1
cat gpurun_out/r2c30_kbench_kbc_variants.txt
