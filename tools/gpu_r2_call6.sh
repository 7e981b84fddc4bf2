#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_tma.py -q > gpurun_out/r2c6_tma_tests.log 2>&1; echo "tma tests rc=$?" >> gpurun_out/r2c6_tma_tests.log
timeout 900 bash tools/kbench.sh run 512 20 > gpurun_out/kbench_r2_c.txt 2>&1; echo "kbench rc=$?" >> gpurun_out/kbench_r2_c.txt
tail -25 gpurun_out/r2c6_tma_tests.log; cat gpurun_out/kbench_r2_c.txt
