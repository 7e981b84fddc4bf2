#!/bin/bash
# round 2, call 26: D2Q9 boundary list fused into the bulk kernel's grid; BGK with Galilean correction; full GPU suite
mkdir -p gpurun_out
{ echo "== cavity, list fused into the bulk kernel (default)"; timeout 600 python tools/cavity_probe.py; echo "== LBMX_NO_FUSED_LIST=1"; LBMX_NO_FUSED_LIST=1 timeout 600 python tools/cavity_probe.py; } > gpurun_out/r2c26_cavity.txt 2>&1
{
for v in q9_f64_fused q9_f32_fused; do timeout 300 ./tools/bin/kb_$v 8192 20 0 0 0 | grep -v "TMA" | tail -1; done
timeout 300 ./tools/bin/kb_q9_f64_fused 1024 500 0 0 1 | grep -v "TMA" | tail -2
for v in bgk_f64 bgkgal_f64; do timeout 300 ./tools/bin/kb_$v 384 20 0 1 0 | grep -v "TMA" | tail -1; done
} > gpurun_out/r2c26_kbench.txt 2>&1
timeout 2400 python -m pytest tests -q -m gpu > gpurun_out/r2c26_gpu_suite.log 2>&1; echo "rc=$?" >> gpurun_out/r2c26_gpu_suite.log
cat gpurun_out/r2c26_cavity.txt gpurun_out/r2c26_kbench.txt; tail -5 gpurun_out/r2c26_gpu_suite.log
