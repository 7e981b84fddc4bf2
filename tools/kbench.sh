#!/bin/bash
# build and (on a GPU box) run the kernel-variant sweep: tools/kbench.sh build | run [size] [iters]
cd "$(dirname "$0")/.."
mkdir -p tools/bin
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo --expt-relaxed-constexpr -ccbin /usr/bin/g++"
# name:defines:macro modes to run (0 none, 1 rho,u every step, 2 MACRO_Mean)
variants=(
 "default::0 1 2"
 "macro_plain:-DLBMX_MACRO_HINT=0:1 2"
 "tma_nohint:-DLBMX_TMA_LD_POLICY=0 -DLBMX_TMA_ST_POLICY=0:0"
 "tma_ldhint:-DLBMX_TMA_LD_POLICY=1 -DLBMX_TMA_ST_POLICY=0:0"
 "tma_sthint:-DLBMX_TMA_LD_POLICY=0 -DLBMX_TMA_ST_POLICY=1:0"
 "tma_occ5:-DLBMX_TMA_MINBLOCKS=5:0"
 "tma_occ3:-DLBMX_TMA_MINBLOCKS=3:0"
 "f32::0 1"
 "f32_macro_plain:-DLBMX_MACRO_HINT=0:1"
)
if [ "$1" = "build" ]; then
  for v in "${variants[@]}"; do
    name="${v%%:*}"; rest="${v#*:}"; defs="${rest%%:*}"
    case $name in f32*) defs="$defs -DKB_REAL=float";; esac
    ( nvcc $FLAGS $defs -DKB_NAME="\"$name\"" tools/kbench.cu -o tools/bin/kb_$name 2> tools/bin/kb_$name.log || echo "FAILED $name" ) &
  done
  wait
  ls tools/bin | grep -v log | wc -l
else
  for v in "${variants[@]}"; do
    name="${v%%:*}"; modes="${v##*:}"
    for m in $modes; do ./tools/bin/kb_$name ${2:-512} ${3:-20} $m; done
  done
fi
