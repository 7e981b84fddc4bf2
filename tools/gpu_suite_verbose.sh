#!/bin/bash
# Run the GPU tests in parallel worker processes with one line per finished test, so that a run cut short by a time limit still
# shows which tests failed.  Usage: tools/gpu_suite_verbose.sh SECONDS [pytest -k expression]
limit=${1:-70}
expr=${2:-"not 1000_steps and not full_size"}
cd "$(dirname "$0")/.."
timeout -s INT "$limit" python -m pytest tests/test_gpu_parity.py tests/test_gpu_physics_and_edges.py tests/test_d3q19.py tests/test_gpu_multi.py tests/test_dropin_solvers.py tests/test_gpu_full_size.py \
	-m gpu -n 8 -v --tb=short -k "$expr" 2>&1 | grep -v "PASSED\|SKIPPED" | tail -150
