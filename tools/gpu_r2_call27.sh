#!/bin/bash
# round 2, call 27: full GPU suite on the final library, smoke(), the driver's bench line
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -q -m gpu > gpurun_out/r2c27_gpu_suite.log 2>&1; echo "rc=$?" >> gpurun_out/r2c27_gpu_suite.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2c27_smoke.log 2>&1; echo "rc=$?" >> gpurun_out/r2c27_smoke.log
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2c27_bench.json 2> gpurun_out/r2c27_bench.err; echo "rc=$?" >> gpurun_out/r2c27_bench.err
timeout 300 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2c27_bench_ref.json 2> gpurun_out/r2c27_bench_ref.err; echo "rc=$?" >> gpurun_out/r2c27_bench_ref.err
tail -4 gpurun_out/r2c27_gpu_suite.log; cat gpurun_out/r2c27_smoke.log; cat gpurun_out/r2c27_bench.json | cut -c1-400; tail -2 gpurun_out/r2c27_bench.err; cat gpurun_out/r2c27_bench_ref.json | cut -c1-300
