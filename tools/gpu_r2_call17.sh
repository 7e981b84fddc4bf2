#!/bin/bash
# round 2, call 17: mixed cold path (call / in line per kernel), D2Q9 choice, MACRO_Mean with L2 prefetch, full GPU suite, solid maps
mkdir -p gpurun_out
{
for v in nocold mix nocold mix; do timeout 300 ./tools/bin/kb_$v 512 20 0 1 0 | grep -v "TMA" | tail -1; done
for v in f32_nocold f32_mix f32_nocold f32_mix; do timeout 300 ./tools/bin/kb_$v 512 20 0 1 0 | grep -v "TMA" | tail -1; done
for v in q19_nocold q19_mix q19_nocold q19_mix; do timeout 300 ./tools/bin/kb_$v 512 20 0 1 0 | grep -v "TMA" | tail -1; done
for v in q9_f64_nocold q9_f64_call q9_f64_inl q9_f64_nocold q9_f64_call q9_f64_inl; do timeout 300 ./tools/bin/kb_$v 8192 20 0 0 0 | grep -v "TMA" | tail -1; done
for v in q9_f32_nocold q9_f32_call q9_f32_inl q9_f32_nocold q9_f32_call q9_f32_inl; do timeout 300 ./tools/bin/kb_$v 8192 20 0 0 0 | grep -v "TMA" | tail -1; done
echo "== MACRO_Mean (macro=2), rho,u every step (macro=1)"
for m in 2 1; do timeout 300 ./tools/bin/kb_mix 512 20 $m 1 0 | grep -v "TMA" | tail -1; done
timeout 300 ./tools/bin/kb_q9_f64_call 8192 20 2 0 0 | grep -v "TMA" | tail -1
} > gpurun_out/r2c17_kbench_mix.txt 2>&1
timeout 2400 python -m pytest tests -q -m gpu -x > gpurun_out/r2c17_gpu_suite.log 2>&1; echo "rc=$?" >> gpurun_out/r2c17_gpu_suite.log
timeout 900 python tools/solid_bench.py --size 384 --streaming AA > gpurun_out/r2c17_solid_aa.jsonl 2> gpurun_out/r2c17_solid_aa.err
timeout 900 python tools/solid_bench.py --size 384 --streaming AB --maps periodic,sphere,wall30,nothing30,nothinghalf > gpurun_out/r2c17_solid_ab.jsonl 2> gpurun_out/r2c17_solid_ab.err
for m in sphere wall30 nothinghalf; do
  timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum -k regex:'k_bulk|k_boundary' --clock-control none --csv --log-file gpurun_out/r2c17_ncu_solid_$m.csv python tools/solid_bench.py --size 384 --maps $m --steps 2 --warmup 2 > /dev/null 2>&1
done
cat gpurun_out/r2c17_kbench_mix.txt; tail -3 gpurun_out/r2c17_gpu_suite.log; cut -c1-330 gpurun_out/r2c17_solid_aa.jsonl gpurun_out/r2c17_solid_ab.jsonl
