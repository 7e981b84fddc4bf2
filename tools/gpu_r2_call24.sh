#!/bin/bash
# round 2, call 24: cells per thread again, now that the A-B kernels lost their face re-load
mkdir -p gpurun_out
{
for v in q19_cpt2 q19_cpt1 f64_cpt2; do timeout 300 ./tools/bin/kb_$v 512 20 0 1 0 | grep -v "TMA" | tail -1; done
for v in q9_f64_cpt1 q9_f64_cpt2 q9_f32_cpt1 q9_f32_cpt2 q9_f32_cpt4; do timeout 300 ./tools/bin/kb_$v 8192 20 0 0 0 | grep -v "TMA" | tail -1; done
for v in q19_cpt2 q19_cpt1; do timeout 300 ./tools/bin/kb_$v 512 20 0 1 0 | grep -v "TMA" | tail -1; done
} > gpurun_out/r2c24_kbench_cpt.txt 2>&1
cat gpurun_out/r2c24_kbench_cpt.txt
