#!/bin/bash
mkdir -p gpurun_out
timeout 900 bash tools/kbench.sh run 512 20 > gpurun_out/kbench_r2_d.txt 2>&1; echo "kbench rc=$?" >> gpurun_out/kbench_r2_d.txt
grep -v "AB " gpurun_out/kbench_r2_d.txt | grep -v TMA
