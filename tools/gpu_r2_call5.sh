#!/bin/bash
mkdir -p gpurun_out
P=./tools/bin/tma_probe3
{
# dtype rank bi bo x style
timeout 30 $P 0 2 32 8 64 0    # the guide's example
timeout 30 $P 0 2 32 8 65 0    # inner coordinate 4 B off a 16 B boundary
timeout 30 $P 0 2 32 8 66 0    # 8 B off
timeout 30 $P 0 2 32 8 68 0    # 16 B: aligned again
timeout 30 $P 0 2 32 1 64 0    # one row
timeout 30 $P 0 2 64 8 64 0    # 256 B inner
timeout 30 $P 0 2 64 1 64 0
timeout 30 $P 1 2 32 8 64 0    # fp32
timeout 30 $P 2 2 16 8 32 0    # fp64
timeout 30 $P 2 2 32 1 32 0    # fp64, 256 B inner, one row
timeout 30 $P 2 2 32 1 33 0    # fp64, 8 B off
timeout 30 $P 2 4 32 1 32 0    # 4-D
timeout 30 $P 2 4 32 1 33 0
timeout 30 $P 0 2 32 8 64 1    # raw mbarrier
timeout 30 $P 2 4 32 1 32 1
timeout 30 $P 2 4 128 1 0 1
} > gpurun_out/r2c5_probe.txt 2>&1
cat gpurun_out/r2c5_probe.txt
