#!/bin/bash
# round 2, call 36: SRT / BGK / KBC fp32 back in the reference's incremental update form (one rounding at the magnitude of f per step): the GPU
# parity tests, the kernels' rates, the cavity of configs[1]; and a sweep over the fp32 cumulant A-A odd kernel (5.0 TB/s against 6.2 even)
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py -q -m gpu > gpurun_out/r2c36_parity.log 2>&1; echo "rc=$?" >> gpurun_out/r2c36_parity.log
{
for v in srt_f64 srt_f32 q19srt_f64 q19srt_f32 bgk_f64 kbcn4_f32; do timeout 300 ./tools/bin/kb_$v 384 20 0 1 1 | grep -v "TMA" | tail -2; done
for v in q9_f64 q9_f32; do timeout 300 ./tools/bin/kb_$v 8192 20 0 0 1 | grep -v "TMA" | tail -2; done
} > gpurun_out/r2c36_kbench_incremental.txt 2>&1
{
for v in f32_default f32_call f32_stream f32_streamx f32_ldcg f32_cpt2 f32_default; do timeout 300 ./tools/bin/kb_$v 512 20 0 1 0 | grep -v "TMA" | tail -1; done
} > gpurun_out/r2c36_kbench_f32_odd.txt 2>&1
timeout 600 python tools/cavity_probe.py > gpurun_out/r2c36_cavity.txt 2>&1
tail -3 gpurun_out/r2c36_parity.log; cat gpurun_out/r2c36_kbench_incremental.txt gpurun_out/r2c36_kbench_f32_odd.txt; head -4 gpurun_out/r2c36_cavity.txt
