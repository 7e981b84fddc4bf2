#!/bin/bash
# round 2, call 10: deferred batches in the host mirror -- invisibility tests, box3d at 512^3, default bench line
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_dropin_solvers.py -x -q -m gpu > gpurun_out/r2c10_dropin.log 2>&1; echo "rc=$?" >> gpurun_out/r2c10_dropin.log
( cd /tmp && timeout 300 $GRAFT_REPO_ROOT/examples/bin/box3d_aa 512 512 512 330 110 ) > gpurun_out/r2c10_box3d.log 2>&1; echo "rc=$?" >> gpurun_out/r2c10_box3d.log
( cd /tmp && LBMX_HOST_BATCH=0 timeout 300 $GRAFT_REPO_ROOT/examples/bin/box3d_aa 512 512 512 120 60 ) > gpurun_out/r2c10_box3d_nobatch.log 2>&1; echo "rc=$?" >> gpurun_out/r2c10_box3d_nobatch.log
timeout 900 python bench.py > gpurun_out/r2c10_bench.json 2> gpurun_out/r2c10_bench.err; echo "rc=$?" >> gpurun_out/r2c10_bench.err
tail -5 gpurun_out/r2c10_dropin.log; grep -E "GLUPS|lbmx:|iterations|rc=" gpurun_out/r2c10_box3d.log gpurun_out/r2c10_box3d_nobatch.log; cat gpurun_out/r2c10_bench.json; tail -3 gpurun_out/r2c10_bench.err
