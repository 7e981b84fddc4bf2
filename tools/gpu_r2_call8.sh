#!/bin/bash
# where do the 3 % between tools/kbench (15.2-15.3 GLUPS) and bench.py (14.5-15.0 GLUPS) go?
mkdir -p gpurun_out
{
./tools/bin/kb_default 512 20 0 | grep -v TMA
for s in smi none nvml; do
  echo "== bench --steps 20 --clock-sampler $s"
  python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extras --clock-sampler $s 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['clocks'], d['e2e']['value'])"
done
for s in smi none nvml; do
  echo "== bench --steps 100 --clock-sampler $s"
  python bench.py --steps 100 --warmup 5 --no-cpu-baseline --no-extras --clock-sampler $s 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['clocks'], d['e2e']['value'])"
done
./tools/bin/kb_default 512 20 0 | grep -v TMA
} > gpurun_out/r2c8_gap.txt 2>&1
cat gpurun_out/r2c8_gap.txt
