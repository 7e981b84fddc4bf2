#!/bin/bash
# round 2, call 12: full GPU suite after "walls in the bulk kernel"; hint variants of the odd step per lattice / precision; obstacle-heavy maps
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -x -q -m gpu > gpurun_out/r2c12_gpu_suite.log 2>&1; echo "rc=$?" >> gpurun_out/r2c12_gpu_suite.log
{
for v in default odd_plain; do timeout 300 ./tools/bin/kb_$v 512 20 0 1 1 | grep -v "TMA\|AB "; done
for v in f32 f32_odd_plain; do timeout 300 ./tools/bin/kb_$v 512 20 0 1 1 | grep -v "TMA\|AB "; done
for v in q19_f32 q19_f32_odd_plain; do timeout 300 ./tools/bin/kb_$v 512 20 0 1 1 | grep -v "TMA\|AB "; done
for v in q9_f64 q9_f64_odd_plain q9_f32 q9_f32_odd_plain; do timeout 300 ./tools/bin/kb_$v 8192 20 0 0 1 | grep -v "TMA\|AB "; done
for v in default odd_plain; do timeout 300 ./tools/bin/kb_$v 512 20 0 1 1 | grep -v "TMA\|AB "; done
} > gpurun_out/r2c12_kbench_hints.txt 2>&1
timeout 900 python tools/solid_bench.py --size 384 --streaming AA > gpurun_out/r2c12_solid_aa.jsonl 2> gpurun_out/r2c12_solid_aa.err
timeout 900 python tools/solid_bench.py --size 384 --streaming AB > gpurun_out/r2c12_solid_ab.jsonl 2> gpurun_out/r2c12_solid_ab.err
for m in periodic sphere wall30blocks wall30 nothing30; do
  timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum -k regex:'k_bulk|k_boundary' --clock-control none --csv --log-file gpurun_out/r2c12_ncu_solid_$m.csv python tools/solid_bench.py --size 384 --maps $m --steps 2 --warmup 2 > /dev/null 2>&1
done
tail -4 gpurun_out/r2c12_gpu_suite.log; cat gpurun_out/r2c12_kbench_hints.txt; cat gpurun_out/r2c12_solid_aa.jsonl gpurun_out/r2c12_solid_ab.jsonl | cut -c1-400
