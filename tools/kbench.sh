#!/bin/bash
# build and (on a GPU box) run the kernel-variant sweep: tools/kbench.sh build | run [size] [iters]
cd "$(dirname "$0")/.."
mkdir -p tools/bin
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo --expt-relaxed-constexpr -ccbin /usr/bin/g++"
variants=(
 "default:"
 "m1:-DLBMX_BULK_MINBLOCKS=1"
 "m3:-DLBMX_BULK_MINBLOCKS=3"
 "m5:-DLBMX_BULK_MINBLOCKS=5"
 "plain:-DLBMX_LD_HINT=0 -DLBMX_ST_HINT=0"
 "cg_wt:-DLBMX_ST_HINT=3"
 "cs_cs:-DLBMX_LD_HINT=1"
 "b64:-DLBMX_BULK_BLOCK=64 -DLBMX_BULK_MINBLOCKS=8"
 "nocollide:-DLBMX_EXP_NOCOLLIDE"
 "noyshift:-DLBMX_EXP_NOYSHIFT"
)
if [ "$1" = "build" ]; then
  for v in "${variants[@]}"; do
    name="${v%%:*}"; defs="${v#*:}"
    ( nvcc $FLAGS $defs -DKB_NAME="\"$name\"" tools/kbench.cu -o tools/bin/kb_$name 2> tools/bin/kb_$name.log || echo "FAILED $name" ) &
  done
  wait
  ls tools/bin | grep -v log | wc -l
else
  for v in "${variants[@]}"; do name="${v%%:*}"; ./tools/bin/kb_$name ${2:-256} ${3:-20}; done
fi
