#!/bin/bash
# build and (on a GPU box) run the kernel-variant sweep: tools/kbench.sh build | run [size] [iters]
cd "$(dirname "$0")/.."
mkdir -p tools/bin
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo --expt-relaxed-constexpr -ccbin /usr/bin/g++"
# name:defines:macro modes to run (0 none, 1 rho,u every step, 2 MACRO_Mean)
variants=(
 "default::0 1 2"
 "odd_ld_plain_y:-DLBMX_LD_HINT_YSHIFT=0:0"
 "odd_ld_plain_all:-DLBMX_LD_HINT_YSHIFT=0 -DLBMX_LD_HINT_ODD=0:0"
 "odd_st_plain_y:-DLBMX_ST_HINT_YSHIFT=0:0"
 "odd_ldst_plain_y:-DLBMX_LD_HINT_YSHIFT=0 -DLBMX_ST_HINT_YSHIFT=0:0"
 "odd_ldst_plain_all:-DLBMX_LD_HINT_YSHIFT=0 -DLBMX_ST_HINT_YSHIFT=0 -DLBMX_LD_HINT_ODD=0 -DLBMX_ST_HINT_ODD=0:0"
 "odd_ld_plain_st_cg_y:-DLBMX_LD_HINT_YSHIFT=0 -DLBMX_ST_HINT_YSHIFT=2:0"
 "odd_ld_cs_y:-DLBMX_LD_HINT_YSHIFT=1:0"
 "f32::0"
 "f32_odd_ld_plain_y:-DLBMX_LD_HINT_YSHIFT=0:0"
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
