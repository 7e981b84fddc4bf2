#!/bin/bash
# build and (on a GPU box) run the kernel-variant sweep: tools/kbench.sh build | run [size] [iters]
cd "$(dirname "$0")/.."
mkdir -p tools/bin
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo --expt-relaxed-constexpr -ccbin /usr/bin/g++"
variants=(
 "default:"
 "f64_cpt2_m2:-DLBMX_BULK_CPT=2 -DLBMX_BULK_MINBLOCKS=2 -DLBMX_BULK_MINBLOCKS_AB=2"
 "f64_cpt2_m3:-DLBMX_BULK_CPT=2 -DLBMX_BULK_MINBLOCKS=3 -DLBMX_BULK_MINBLOCKS_AB=3"
 "f32_cpt1:-DKB_REAL=float -DLBMX_BULK_CPT=1"
 "f32_cpt2:-DKB_REAL=float -DLBMX_BULK_CPT=2"
 "f32_cpt2_m3:-DKB_REAL=float -DLBMX_BULK_CPT=2 -DLBMX_BULK_MINBLOCKS=3 -DLBMX_BULK_MINBLOCKS_AB=3"
 "f32_cpt4_m2:-DKB_REAL=float -DLBMX_BULK_CPT=4 -DLBMX_BULK_MINBLOCKS=2 -DLBMX_BULK_MINBLOCKS_AB=2"
 "f32_cpt2_b256:-DKB_REAL=float -DLBMX_BULK_CPT=2 -DLBMX_BULK_BLOCK=256 -DLBMX_BULK_MINBLOCKS=2 -DLBMX_BULK_MINBLOCKS_AB=2"
 "q9_f64_cpt1:-DKB_LAT=D2Q9 -DKB_KIND=K_SRT -DLBMX_BULK_CPT=1"
 "q9_f64_cpt2:-DKB_LAT=D2Q9 -DKB_KIND=K_SRT -DLBMX_BULK_CPT=2"
 "q9_f64_cpt4:-DKB_LAT=D2Q9 -DKB_KIND=K_SRT -DLBMX_BULK_CPT=4"
 "q9_f32_cpt1:-DKB_LAT=D2Q9 -DKB_KIND=K_SRT -DKB_REAL=float -DLBMX_BULK_CPT=1"
 "q9_f32_cpt2:-DKB_LAT=D2Q9 -DKB_KIND=K_SRT -DKB_REAL=float -DLBMX_BULK_CPT=2"
 "q9_f32_cpt4:-DKB_LAT=D2Q9 -DKB_KIND=K_SRT -DKB_REAL=float -DLBMX_BULK_CPT=4"
 "q9_f32_cpt8:-DKB_LAT=D2Q9 -DKB_KIND=K_SRT -DKB_REAL=float -DLBMX_BULK_CPT=8"
 "q9_f32_cpt4_b256:-DKB_LAT=D2Q9 -DKB_KIND=K_SRT -DKB_REAL=float -DLBMX_BULK_CPT=4 -DLBMX_BULK_BLOCK=256 -DLBMX_BULK_MINBLOCKS=2 -DLBMX_BULK_MINBLOCKS_AB=2"
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
