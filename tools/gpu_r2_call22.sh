#!/bin/bash
# round 2, call 22: A-B kernels without the face re-load (face fluid cells in the boundary list) + whole-sector stores
mkdir -p gpurun_out
{
for v in ws2 f32_ws2 q19_nows2 q19_ws2 q19srt_ws2; do timeout 300 ./tools/bin/kb_$v 512 20 0 1 0 | grep -v "TMA" | tail -1; done
for v in q9_f64_ws2 q9_f32_ws2; do timeout 300 ./tools/bin/kb_$v 8192 20 0 0 0 | grep -v "TMA" | tail -1; done
for v in ws2 f32_ws2 q19_nows2 q19_ws2; do timeout 300 ./tools/bin/kb_$v 512 20 0 1 0 | grep -v "TMA" | tail -1; done
} > gpurun_out/r2c22_kbench_ws2.txt 2>&1
timeout 900 python tools/solid_bench.py --size 384 --streaming AB --maps periodic,fluid,shell,columns,duct,sim1,sphere,wall30 > gpurun_out/r2c22_solid_ab.jsonl 2>&1
timeout 2400 python -m pytest tests -q -m gpu > gpurun_out/r2c22_gpu_suite.log 2>&1; echo "rc=$?" >> gpurun_out/r2c22_gpu_suite.log
{ echo "== cavity (PDL chain)"; timeout 600 python tools/cavity_probe.py; } > gpurun_out/r2c22_cavity.txt 2>&1
cat gpurun_out/r2c22_kbench_ws2.txt; tail -3 gpurun_out/r2c22_gpu_suite.log; cat gpurun_out/r2c22_cavity.txt
python - <<'PY'
import json
for ln in open("gpurun_out/r2c22_solid_ab.jsonl"):
    if ln.startswith("{"):
        d=json.loads(ln); print(d["streaming"], d["map"], "%.3f ms"%d["ms_per_step"], "%.0f GB/s alg"%d["GBs_algorithmic"], "list", d["boundary_list_cells"])
    else: print(ln.strip()[:200])
PY
