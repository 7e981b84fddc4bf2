#!/bin/bash
# round 2, call 34: which A-B bulk kernels pay for the whole-sector path (obstacle / inert lanes store with their warp) on a map without
# obstacles: D2Q9 and D3Q19 kernels, fp32 D3Q27, with the path (default) and without (ws0)
mkdir -p gpurun_out
{
for v in q9_f64 q9_f64_ws0 q9_f32 q9_f32_ws0 q9clbm_f64 q9clbm_f64_ws0 q9clbm_f32 q9clbm_f32_ws0; do timeout 300 ./tools/bin/kb_$v 8192 20 0 0 0 | grep -v "TMA" | tail -1; done
for v in q19srt_f32 q19srt_f32_ws0 q19mrt_f32 q19mrt_f32_ws0; do timeout 300 ./tools/bin/kb_$v 512 20 0 1 0 | grep -v "TMA" | tail -1; done
for v in cum_f32 cum_f32_ws0 srt_f32_ws0; do timeout 300 ./tools/bin/kb_$v 512 20 0 1 0 | grep -v "TMA" | tail -1; done
} > gpurun_out/r2c34_kbench_ws.txt 2>&1
cat gpurun_out/r2c34_kbench_ws.txt
