#!/bin/bash
# build and (on a GPU box) run the kernel-variant sweep: tools/kbench.sh build | run [size] [iters]
cd "$(dirname "$0")/.."
mkdir -p tools/bin
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo --expt-relaxed-constexpr -ccbin /usr/bin/g++"
variants=(
 "default:"
 "noyshift:-DLBMX_EXP_NOYSHIFT"
 "yshift4:-DLBMX_EXP_YSHIFT4"
)
if [ "$1" = "build" ]; then
  for v in "${variants[@]}"; do
    name="${v%%:*}"; defs="${v#*:}"
    ( nvcc $FLAGS $defs -DKB_NAME="\"$name\"" tools/kbench.cu -o tools/bin/kb_$name 2> tools/bin/kb_$name.log || echo "FAILED $name" ) &
  done
  wait
  ls tools/bin | grep -v log | wc -l
else
  for v in "${variants[@]}"; do name="${v%%:*}"; sz=${2:-256}; case $name in q9*) sz=${4:-8192};; esac; ./tools/bin/kb_$name $sz ${3:-20}; done
fi
