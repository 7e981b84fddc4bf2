#!/bin/bash
# round 2, call 16: cold path as an out-of-line function that reloads (hot path size / registers as before?)
mkdir -p gpurun_out
{
for v in nocold call nocold call; do timeout 300 ./tools/bin/kb_$v 512 20 0 1 0 | grep -v "TMA" | tail -1; done
for v in f32_nocold f32_call f32_nocold f32_call; do timeout 300 ./tools/bin/kb_$v 512 20 0 1 0 | grep -v "TMA" | tail -1; done
for v in q19_orig q19_nocold q19_call q19_orig q19_call; do timeout 300 ./tools/bin/kb_$v 512 20 0 1 0 | grep -v "TMA" | tail -1; done
} > gpurun_out/r2c16_kbench_call.txt 2>&1
cat gpurun_out/r2c16_kbench_call.txt
