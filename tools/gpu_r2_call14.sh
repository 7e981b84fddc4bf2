#!/bin/bash
# round 2, call 14: cold path with an out-of-line macro writer; L2 residency of small lattices (cache hints at 1024^2)
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -q -m gpu -x > gpurun_out/r2c14_gpu_suite.log 2>&1; echo "rc=$?" >> gpurun_out/r2c14_gpu_suite.log
{
for v in nocold cold2 nocold cold2; do timeout 300 ./tools/bin/kb_$v 512 20 0 1 1 | grep -v "TMA"| tail -2; done
for v in nocold cold2 nocold cold2; do timeout 300 ./tools/bin/kb_f32_$v 512 20 0 1 1 | grep -v "TMA" | tail -2; done
} > gpurun_out/r2c14_kbench_cold2.txt 2>&1
{
for S in 512 1024 2048; do for v in q9_f64 q9_f64_plain q9_f64_cg_plain q9_f64_plain_cs q9_f64_cg_cg; do timeout 120 ./tools/bin/kb_$v $S 500 0 0 1 | grep -v "TMA" | tail -2; done; done
} > gpurun_out/r2c14_kbench_q9_l2.txt 2>&1
tail -4 gpurun_out/r2c14_gpu_suite.log; cat gpurun_out/r2c14_kbench_cold2.txt gpurun_out/r2c14_kbench_q9_l2.txt
