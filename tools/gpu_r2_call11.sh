#!/bin/bash
# round 2, call 11: deferral fixed (stat_counter); where do the 3 % between per-kernel timing and lbmx_step() go: data or chaining?
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_dropin_solvers.py -x -q -m gpu > gpurun_out/r2c11_dropin.log 2>&1; echo "rc=$?" >> gpurun_out/r2c11_dropin.log
( cd /tmp && timeout 300 $GRAFT_REPO_ROOT/examples/bin/box3d_aa 512 512 512 330 110 ) > gpurun_out/r2c11_box3d.log 2>&1; echo "rc=$?" >> gpurun_out/r2c11_box3d.log
{
for f in 0 1; do timeout 300 ./tools/bin/kb_default 512 20 0 $f 1 | grep -v TMA; done
for f in 0 1; do timeout 300 ./tools/bin/kb_odd_ldst_plain_all 512 20 0 $f 1 | grep -v TMA; done
timeout 300 ./tools/bin/kb_default 512 60 0 1 1 | grep -v TMA
} > gpurun_out/r2c11_kbench_field_chain.txt 2>&1
tail -5 gpurun_out/r2c11_dropin.log; grep -E "GLUPS|lbmx:|iterations|rc=" gpurun_out/r2c11_box3d.log; cat gpurun_out/r2c11_kbench_field_chain.txt
