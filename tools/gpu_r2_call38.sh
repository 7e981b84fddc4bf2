#!/bin/bash
# round 2, call 38: final library -- full GPU suite, smoke(), the kernel table (kbench) and the kernel-family table through the engine, the
# cavity of configs[1], the driver's bench line and the reference arm
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -q -m gpu > gpurun_out/r2c38_gpu_suite.log 2>&1; echo "rc=$?" >> gpurun_out/r2c38_gpu_suite.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2c38_smoke.log 2>&1; echo "rc=$?" >> gpurun_out/r2c38_smoke.log
{
timeout 300 ./tools/bin/kb_final 512 20 0 1 1 | grep -v "TMA" | tail -2
for m in 1 2; do timeout 300 ./tools/bin/kb_final 512 20 $m 1 0 | grep -v "TMA" | tail -1; done
timeout 300 ./tools/bin/kb_f32_final 512 20 0 1 1 | grep -v "TMA" | tail -2
timeout 300 ./tools/bin/kb_q19_final 512 20 0 1 1 | grep -v "TMA" | tail -2
timeout 300 ./tools/bin/kb_q9_f64_final 8192 20 0 0 1 | grep -v "TMA" | tail -2
timeout 300 ./tools/bin/kb_q9_f64_final 8192 20 2 0 0 | grep -v "TMA" | tail -1
timeout 300 ./tools/bin/kb_q9_f32 8192 20 0 0 1 | grep -v "TMA" | tail -2
for v in srt_f64 srt_f32 q19srt_f64 q19srt_f32 bgk_f64 kbcn4_f64 kbcn4_f32 clbm_f64 clbm_f32 cum2017aa_f64 cum2017aa_f32; do timeout 300 ./tools/bin/kb_$v 384 20 0 1 1 | grep -v "TMA" | tail -2; done
} > gpurun_out/r2c38_kbench_final.txt 2>&1
timeout 600 python tools/cavity_probe.py > gpurun_out/r2c38_cavity.txt 2>&1
timeout 1500 python tools/bench_matrix.py --out gpurun_out/r2c38_bench_matrix.md > gpurun_out/r2c38_bench_matrix.log 2>&1; echo "rc=$?" >> gpurun_out/r2c38_bench_matrix.log
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2c38_bench.json 2> gpurun_out/r2c38_bench.err; echo "rc=$?" >> gpurun_out/r2c38_bench.err
timeout 300 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2c38_bench_ref.json 2> gpurun_out/r2c38_bench_ref.err; echo "rc=$?" >> gpurun_out/r2c38_bench_ref.err
tail -4 gpurun_out/r2c38_gpu_suite.log; cat gpurun_out/r2c38_smoke.log; cat gpurun_out/r2c38_kbench_final.txt; head -4 gpurun_out/r2c38_cavity.txt; cat gpurun_out/r2c38_bench.json | cut -c1-300; cat gpurun_out/r2c38_bench_ref.json | cut -c1-150
