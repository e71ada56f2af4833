#!/bin/bash
# Where the embed kernel's time goes, measured in place: builds of libgswm with ONE ingredient switched off each
# (wrong latents, timing only), timed by tools/kbench.py next to the product build.
#   bash tools/whatif.sh build      (here: nvcc cross-compiles)      bash tools/whatif.sh run   (under gpurun)
set -e
cd "$(dirname "$0")/.."
V=build_variants
declare -A F=( [product]="" [philox1]="-DGSWM_PHILOX_ROUNDS=1" [philox4]="-DGSWM_PHILOX_ROUNDS=4" [philox10]="-DGSWM_PHILOX_ROUNDS=10"
               [nopoly]="-DGSWM_WHATIF_POLY_SKIP=7" [notail]="-DGSWM_WHATIF_NOTAIL" [nosign]="-DGSWM_WHATIF_NOSIGN"
               [nostore]="-DGSWM_WHATIF_NOSTORE" [bare]="-DGSWM_PHILOX_ROUNDS=1 -DGSWM_WHATIF_POLY_SKIP=7 -DGSWM_WHATIF_NOTAIL -DGSWM_WHATIF_NOSIGN" )
ORDER="product philox10 philox4 philox1 nopoly notail nosign nostore bare"
if [ "$1" = build ]; then
  mkdir -p $V
  for n in $ORDER; do
    nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -shared -Iinclude ${F[$n]} \
         -o $V/whatif_$n.so a-watermark-for-diffusion-models_b200/csrc/gswm_kernels.cu a-watermark-for-diffusion-models_b200/csrc/gswm_pipe.cu &
  done
  wait
  ls $V
else
  for n in $ORDER; do
    KB_ONLY_EMBED=1 python tools/kbench.py $V/whatif_$n.so 2>&1 | tail -1
  done
fi
