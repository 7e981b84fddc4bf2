#!/bin/bash
# round 2, call 19: where to put the L2 prefetch of the MACRO_Mean sums so that the kernels that do not use it keep their speed
mkdir -p gpurun_out
{
for P in P0 P1 P2 P3 P4; do timeout 300 ./tools/bin/kb_${P}_f64 512 20 0 1 0 | grep -v "TMA" | tail -1; done
for P in P0 P1 P2 P3 P4; do timeout 300 ./tools/bin/kb_${P}_f64 512 20 2 1 0 | grep -v "TMA" | tail -1; done
for P in P0 P1 P2 P3 P4; do timeout 300 ./tools/bin/kb_${P}_f32 512 20 0 1 0 | grep -v "TMA" | tail -1; done
for P in P0 P1 P2 P3 P4; do timeout 300 ./tools/bin/kb_${P}_q19 512 20 0 1 0 | grep -v "TMA" | tail -1; done
for P in P0 P1 P2 P3 P4; do timeout 300 ./tools/bin/kb_${P}_kbc 384 20 0 1 0 | grep -v "TMA" | tail -1; done
for P in P0 P1; do timeout 300 ./tools/bin/kb_${P}_f32 512 20 0 1 0 | grep -v "TMA" | tail -1; done
} > gpurun_out/r2c19_kbench_prefetch.txt 2>&1
# A-B maps with inflow / outflow planes: what does the boundary-list kernel cost?
for m in sphere duct; do
  timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,launch__registers_per_thread -k regex:'k_bulk|k_boundary' --clock-control none --csv --log-file gpurun_out/r2c19_ncu_ab_$m.csv python tools/solid_bench.py --size 384 --maps $m --steps 2 --warmup 2 --streaming AB > /dev/null 2>&1
done
cat gpurun_out/r2c19_kbench_prefetch.txt
