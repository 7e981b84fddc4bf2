#!/bin/bash
# round 2, call 32: D2Q9 SRT without the three constant divisions per cell (weights as constants in default arithmetic), mean-collecting macro
# flavours out of line in the D2Q9 bulk kernels (minl = in line as before); D3Q19 SRT fp64 A-B without the whole-sector path
mkdir -p gpurun_out
{
for v in q9_f64 q9_f64_minl q9_f64; do timeout 300 ./tools/bin/kb_$v 8192 20 0 0 1 | grep -v "TMA" | tail -2; done
timeout 300 ./tools/bin/kb_q9_f32 8192 20 0 0 1 | grep -v "TMA" | tail -2
timeout 300 ./tools/bin/kb_q9_f64 8192 20 2 0 0 | grep -v "TMA" | tail -1
timeout 300 ./tools/bin/kb_q9_f64_minl 8192 20 2 0 0 | grep -v "TMA" | tail -1
timeout 300 ./tools/bin/kb_q19srt_f64 384 20 0 1 1 | grep -v "TMA" | tail -2
} > gpurun_out/r2c32_kbench_q9.txt 2>&1
cat gpurun_out/r2c32_kbench_q9.txt
