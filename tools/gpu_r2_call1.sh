#!/bin/bash
# round 2, GPU call 1: TMA kernel correctness, kernel-variant sweep, parity suite with the TMA kernels switched on
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > gpurun_out/r2c1_smi.txt 2>&1
timeout 600 python -m pytest tests/test_gpu_tma.py -x -q > gpurun_out/r2c1_tma_tests.log 2>&1; echo "tma tests rc=$?" >> gpurun_out/r2c1_tma_tests.log
timeout 900 bash tools/kbench.sh run 512 20 > gpurun_out/kbench_r2_a.txt 2>&1; echo "kbench rc=$?" >> gpurun_out/kbench_r2_a.txt
LBMX_TMA=both timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_multi.py tests/test_gpu_full_size.py -x -q -m gpu > gpurun_out/r2c1_parity_tma.log 2>&1; echo "rc=$?" >> gpurun_out/r2c1_parity_tma.log
tail -3 gpurun_out/r2c1_tma_tests.log gpurun_out/r2c1_parity_tma.log; cat gpurun_out/kbench_r2_a.txt
