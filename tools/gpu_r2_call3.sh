#!/bin/bash
mkdir -p gpurun_out
{
nvidia-smi -q | grep -i -E "mig mode|current|virtualization|Product Name|Driver Version|CUDA Version|Compute Mode|Confidential" | head -20
for s in 0 1 2 3 4 5 6 7; do timeout 60 ./tools/bin/tma_probe $s; echo "rc=$?"; done
} > gpurun_out/r2c3_probe.txt 2>&1
cat gpurun_out/r2c3_probe.txt
