#!/bin/bash
# round 2, call 13: walls / inert cells in the bulk kernel (cold path), PDL chains on small lattices, odd-step hints
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -q -m gpu > gpurun_out/r2c13_gpu_suite.log 2>&1; echo "rc=$?" >> gpurun_out/r2c13_gpu_suite.log
{
for v in nocold cold hotbranch; do timeout 300 ./tools/bin/kb_$v 512 20 0 1 1 | grep -v "TMA"| tail -2; done
for v in nocold cold hotbranch; do timeout 300 ./tools/bin/kb_f32_$v 512 20 0 1 1 | grep -v "TMA" | tail -2; done
for v in nocold cold; do timeout 300 ./tools/bin/kb_$v 512 20 0 1 1 | grep -v "TMA"| tail -2; done
} > gpurun_out/r2c13_kbench_cold.txt 2>&1
{ echo "== PDL chain (default)"; timeout 600 python tools/cavity_probe.py; echo "== LBMX_NO_PDL=1"; LBMX_NO_PDL=1 timeout 600 python tools/cavity_probe.py; } > gpurun_out/r2c13_cavity.txt 2>&1
timeout 900 python tools/solid_bench.py --size 384 --streaming AA > gpurun_out/r2c13_solid_aa.jsonl 2> gpurun_out/r2c13_solid_aa.err
timeout 900 python tools/solid_bench.py --size 384 --streaming AB --maps periodic,sphere,wall30,nothing30,nothinghalf > gpurun_out/r2c13_solid_ab.jsonl 2> gpurun_out/r2c13_solid_ab.err
tail -8 gpurun_out/r2c13_gpu_suite.log; cat gpurun_out/r2c13_kbench_cold.txt gpurun_out/r2c13_cavity.txt; cut -c1-330 gpurun_out/r2c13_solid_aa.jsonl gpurun_out/r2c13_solid_ab.jsonl
