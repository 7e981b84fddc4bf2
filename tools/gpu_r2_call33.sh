#!/bin/bash
# round 2, call 33: SRT in its default-arithmetic form (equilibrium never stored, source bracket as per-axis sums; D2Q9: shared 3 c.u) on
# D3Q27 / D3Q19 / D2Q9, 3 against 4 CTAs per SM in fp64; the cavity of BASELINE configs[1] (1024^2 D2Q9 fp64) and its neighbours
mkdir -p gpurun_out
{
for v in srt_f64 srt_f64_mb4 srt_f32; do timeout 300 ./tools/bin/kb_$v 384 20 0 1 1 | grep -v "TMA" | tail -2; done
for v in q19srt_f64 q19srt_f64_mb4 q19srt_f32; do timeout 300 ./tools/bin/kb_$v 384 20 0 1 1 | grep -v "TMA" | tail -2; done
for v in q9_f64 q9_f32; do timeout 300 ./tools/bin/kb_$v 8192 20 0 0 1 | grep -v "TMA" | tail -2; done
} > gpurun_out/r2c33_kbench_srt.txt 2>&1
timeout 600 python tools/cavity_probe.py > gpurun_out/r2c33_cavity.txt 2>&1
cat gpurun_out/r2c33_kbench_srt.txt gpurun_out/r2c33_cavity.txt
