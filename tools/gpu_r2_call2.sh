#!/bin/bash
mkdir -p gpurun_out
{
for s in 1 2 3 4 5; do timeout 60 ./tools/bin/tma_probe $s; echo "rc=$?"; done
timeout 120 ./tools/bin/kb_diag 64 2 0; echo "kb_diag rc=$?"
timeout 300 compute-sanitizer --tool memcheck ./tools/bin/kb_diag 32 1 0 2>&1 | head -60
} > gpurun_out/r2c2_probe.txt 2>&1
timeout 900 bash tools/kbench.sh run 512 20 > gpurun_out/kbench_r2_b.txt 2>&1
cat gpurun_out/r2c2_probe.txt; cat gpurun_out/kbench_r2_b.txt
