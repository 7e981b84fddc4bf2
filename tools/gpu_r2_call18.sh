#!/bin/bash
# round 2, call 18: final kernel configuration -- kbench table, full GPU suite, the default bench line, ncu launch list and full capture
mkdir -p gpurun_out
{
timeout 300 ./tools/bin/kb_final 512 20 0 1 1 | grep -v "TMA" | tail -2
for m in 1 2; do timeout 300 ./tools/bin/kb_final 512 20 $m 1 0 | grep -v "TMA" | tail -1; done
timeout 300 ./tools/bin/kb_f32_final 512 20 0 1 0 | grep -v "TMA" | tail -1
timeout 300 ./tools/bin/kb_q19_final 512 20 0 1 0 | grep -v "TMA" | tail -1
timeout 300 ./tools/bin/kb_q9_f64_final 8192 20 0 0 0 | grep -v "TMA" | tail -1
timeout 300 ./tools/bin/kb_q9_f64_final 8192 20 2 0 0 | grep -v "TMA" | tail -1
for v in kbcn4_f64 kbcn4_f32 clbm_f64 clbm_f32; do timeout 300 ./tools/bin/kb_$v 384 20 0 1 0 | grep -v "TMA" | tail -1; done
} > gpurun_out/r2c18_kbench_final.txt 2>&1
timeout 2400 python -m pytest tests -q -m gpu > gpurun_out/r2c18_gpu_suite.log 2>&1; echo "rc=$?" >> gpurun_out/r2c18_gpu_suite.log
timeout 1500 python bench.py > gpurun_out/r2c18_bench.json 2> gpurun_out/r2c18_bench.err; echo "rc=$?" >> gpurun_out/r2c18_bench.err
timeout 600 python bench.py --steps 20 --warmup 5 --no-extras --no-cpu-baseline > gpurun_out/r2c18_bench_20.json 2> gpurun_out/r2c18_bench_20.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2c18_launches.csv python bench.py --steps 20 --warmup 5 --no-extras --no-cpu-baseline > gpurun_out/r2c18_ncu_bench.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_bulk -c 2 -o gpurun_out/r2c18_kbulk_512 -f ./tools/bin/kb_final 512 1 > gpurun_out/r2c18_ncu_full.log 2>&1
ncu -i gpurun_out/r2c18_kbulk_512.ncu-rep --page raw --csv > gpurun_out/r2c18_kbulk_512_raw.csv 2>/dev/null
ls -la gpurun_out/ | tail -12
cat gpurun_out/r2c18_kbench_final.txt; tail -3 gpurun_out/r2c18_gpu_suite.log; cat gpurun_out/r2c18_bench.json; tail -4 gpurun_out/r2c18_bench.err; cat gpurun_out/r2c18_bench_20.json | cut -c1-300
