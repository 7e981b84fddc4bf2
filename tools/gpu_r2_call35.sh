#!/bin/bash
# round 2, call 35: the library with the reorganised operators -- full GPU suite, smoke(), kernel-family table through the engine, the driver's
# bench line and reference arm, ncu launch list of the bench and full captures of the CLBM / KBC / cumulant bulk kernels
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -q -m gpu > gpurun_out/r2c35_gpu_suite.log 2>&1; echo "rc=$?" >> gpurun_out/r2c35_gpu_suite.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2c35_smoke.log 2>&1; echo "rc=$?" >> gpurun_out/r2c35_smoke.log
timeout 1500 python tools/bench_matrix.py --out gpurun_out/r2c35_bench_matrix.md > gpurun_out/r2c35_bench_matrix.log 2>&1; echo "rc=$?" >> gpurun_out/r2c35_bench_matrix.log
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2c35_bench.json 2> gpurun_out/r2c35_bench.err; echo "rc=$?" >> gpurun_out/r2c35_bench.err
timeout 300 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2c35_bench_ref.json 2> gpurun_out/r2c35_bench_ref.err; echo "rc=$?" >> gpurun_out/r2c35_bench_ref.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2c35_launches.csv python bench.py --steps 20 --warmup 5 --no-extras --no-cpu-baseline > gpurun_out/r2c35_ncu_bench.log 2>&1
for k in clbm_f64 kbcn4_f64; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_bulk -c 3 -o gpurun_out/r2c35_$k -f ./tools/bin/kb_$k 384 1 > gpurun_out/r2c35_ncu_$k.log 2>&1
  ncu -i gpurun_out/r2c35_$k.ncu-rep --page raw --csv > gpurun_out/r2c35_${k}_raw.csv 2>/dev/null
done
ls -la gpurun_out/ | tail -14
tail -4 gpurun_out/r2c35_gpu_suite.log; cat gpurun_out/r2c35_smoke.log; cat gpurun_out/r2c35_bench.json | cut -c1-400; tail -3 gpurun_out/r2c35_bench.err; cat gpurun_out/r2c35_bench_ref.json | cut -c1-200; tail -75 gpurun_out/r2c35_bench_matrix.log
