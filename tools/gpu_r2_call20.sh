#!/bin/bash
# round 2, call 20: why is the A-B bulk kernel 15 % slower on duct-like maps?  (isolating map features)
mkdir -p gpurun_out
timeout 900 python tools/solid_bench.py --size 384 --streaming AB --maps periodic,fluid,shell,ring,plane,columns,duct > gpurun_out/r2c20_solid_ab_features.jsonl 2>&1
timeout 900 python tools/solid_bench.py --size 384 --streaming AA --maps periodic,shell,ring,plane,columns,duct > gpurun_out/r2c20_solid_aa_features.jsonl 2>&1
python - <<'PY'
import json
for f in ("gpurun_out/r2c20_solid_ab_features.jsonl","gpurun_out/r2c20_solid_aa_features.jsonl"):
    for ln in open(f):
        if ln.startswith("{"):
            d=json.loads(ln); print(d["streaming"], d["map"], "%.3f ms"%d["ms_per_step"], "list", d["boundary_list_cells"])
        else: print(ln.strip()[:200])
PY
