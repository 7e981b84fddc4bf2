#!/bin/bash
# round 2, call 23: fp32 D3Q27 A-B variants; boundary-list kernel at 4 CTAs per SM
mkdir -p gpurun_out
{
for v in f32_ws2 f32_nows2 f32_ws2_cpt1 f32_ws2_mb4 f32_ws2 f32_nows2; do timeout 300 ./tools/bin/kb_$v 512 20 0 1 0 | grep -v "TMA" | tail -1; done
} > gpurun_out/r2c23_kbench_f32ab.txt 2>&1
timeout 900 python tools/solid_bench.py --size 384 --streaming AB --maps periodic,fluid,sim1,sphere,wall30 > gpurun_out/r2c23_solid_ab.jsonl 2>&1
timeout 900 python tools/solid_bench.py --size 384 --streaming AA --maps periodic,sphere,wall30 > gpurun_out/r2c23_solid_aa.jsonl 2>&1
{ echo "== cavity (boundary kernel at 4 CTAs/SM)"; timeout 600 python tools/cavity_probe.py; } > gpurun_out/r2c23_cavity.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_physics_and_edges.py -q -m gpu -x > gpurun_out/r2c23_parity.log 2>&1; echo "rc=$?" >> gpurun_out/r2c23_parity.log
cat gpurun_out/r2c23_kbench_f32ab.txt gpurun_out/r2c23_cavity.txt; tail -3 gpurun_out/r2c23_parity.log
python - <<'PY'
import json
for f in ("gpurun_out/r2c23_solid_ab.jsonl","gpurun_out/r2c23_solid_aa.jsonl"):
    for ln in open(f):
        if ln.startswith("{"):
            d=json.loads(ln); print(d["streaming"], d["map"], "%.3f ms"%d["ms_per_step"], "%.0f GB/s alg"%d["GBs_algorithmic"], "list", d["boundary_list_cells"])
        else: print(ln.strip()[:200])
PY
