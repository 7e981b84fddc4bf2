#!/bin/bash
# round 2, call 21: A-B whole-sector stores (obstacle / inert lanes store with their warp)
mkdir -p gpurun_out
{
for v in nows ws f32_nows f32_ws q19_nows q19_ws; do timeout 300 ./tools/bin/kb_$v 512 20 0 1 0 | grep -v "TMA" | tail -1; done
for v in q9_f64_nows q9_f64_ws; do timeout 300 ./tools/bin/kb_$v 8192 20 0 0 0 | grep -v "TMA" | tail -1; done
for v in nows ws f32_nows f32_ws q19_nows q19_ws; do timeout 300 ./tools/bin/kb_$v 512 20 0 1 0 | grep -v "TMA" | tail -1; done
} > gpurun_out/r2c21_kbench_ws.txt 2>&1
timeout 900 python tools/solid_bench.py --size 384 --streaming AB --maps periodic,fluid,shell,ring,plane,columns,duct,sim1,sphere,wall30,nothing30,nothinghalf > gpurun_out/r2c21_solid_ab.jsonl 2>&1
timeout 900 python tools/solid_bench.py --size 384 --streaming AA --maps periodic,duct,sim1,sphere,wall30,nothing30,nothinghalf > gpurun_out/r2c21_solid_aa.jsonl 2>&1
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_physics_and_edges.py -q -m gpu -x > gpurun_out/r2c21_parity.log 2>&1; echo "rc=$?" >> gpurun_out/r2c21_parity.log
cat gpurun_out/r2c21_kbench_ws.txt; tail -3 gpurun_out/r2c21_parity.log
python - <<'PY'
import json
for f in ("gpurun_out/r2c21_solid_ab.jsonl","gpurun_out/r2c21_solid_aa.jsonl"):
    for ln in open(f):
        if ln.startswith("{"):
            d=json.loads(ln); print(d["streaming"], d["map"], "%.3f ms"%d["ms_per_step"], "%.0f GB/s alg"%d["GBs_algorithmic"], "list", d["boundary_list_cells"])
        else: print(ln.strip()[:200])
PY
