#!/bin/bash
# round 2, call 15: what slows the fp32 A-A even kernel down (bisect), KBC occupancy variants
mkdir -p gpurun_out
{
for v in f32_orig1 f32_nocold1 f32_cur f32_nopdl f32_noinert f32_orig1 f32_cur; do timeout 300 ./tools/bin/kb_$v 512 20 0 1 0 | grep -v "TMA" | tail -1; done
for v in q19_orig q19_nocold q19_cur q19_orig q19_cur; do timeout 300 ./tools/bin/kb_$v 512 20 0 1 0 | grep -v "TMA" | tail -1; done
} > gpurun_out/r2c15_kbench_f32_bisect.txt 2>&1
{
for v in kbc_f64_old_mb2 kbc_f64_new_mb2 kbc_f64_new_mb3 kbc_f32_old_mb3 kbc_f32_new_mb3 kbc_f32_new_mb4; do timeout 300 ./tools/bin/kb_$v 384 20 0 1 0 | grep -v "TMA" | tail -1; done
} > gpurun_out/r2c15_kbench_kbc.txt 2>&1
cat gpurun_out/r2c15_kbench_f32_bisect.txt gpurun_out/r2c15_kbench_kbc.txt
