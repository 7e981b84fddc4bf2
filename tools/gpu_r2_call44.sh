#!/bin/bash
# round 2, call 44: smoke() with the two added cases
mkdir -p gpurun_out
timeout 100 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2c44_smoke.log 2>&1; echo "rc=$?" >> gpurun_out/r2c44_smoke.log
cat gpurun_out/r2c44_smoke.log
