#!/bin/bash
mkdir -p gpurun_out
{
env | grep -i -E "cuda|inject|preload|nsight|cupti" 
cat /proc/self/maps | grep -i -E "inject|cupti" | head -3
for v in 0 1 2; do timeout 60 ./tools/bin/tma_probe2 $v; echo "rc=$?"; done
timeout 60 ./tools/bin/tma_probe2_sm100 0; echo "sm100 rc=$?"
timeout 60 ./tools/bin/tma_probe2_ptx90 0; echo "ptx90 (JIT by the driver) rc=$?"
python - <<'PY'
import torch
a=torch.randn(4096,4096,device='cuda',dtype=torch.bfloat16); b=a@a; torch.cuda.synchronize(); print('torch matmul ok', float(b.float().abs().mean()))
PY
} > gpurun_out/r2c4_probe.txt 2>&1
cat gpurun_out/r2c4_probe.txt
