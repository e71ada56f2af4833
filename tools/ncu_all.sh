#!/bin/bash
# ncu --set full capture of the LAST of three launches of every kernel variant (one kernel per ncu run), reports into
# gpurun_out/<tag>_<target>.ncu-rep.  Usage (under gpurun): bash tools/ncu_all.sh r01d [targets...]
tag=${1:-r01d}; shift
targets=${@:-embed_shared embed_per_latent embed_injected extract_f32 extract_f16 extract_per_latent keystream}
declare -A K=( [embed_shared]=embed_kernel [embed_per_latent]=embed_kernel [embed_injected]=embed_injected_kernel
               [extract_f32]=extract_kernel [extract_f16]=extract_kernel [extract_bf16]=extract_kernel
               [extract_per_latent]=extract_kernel [keystream]=chacha20_keystream_kernel [mt19937]=mt19937_kernel )
mkdir -p gpurun_out
for t in $targets; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:${K[$t]} -s 2 -c 1 -f \
      -o gpurun_out/${tag}_${t} python tools/ncu_targets.py $t > gpurun_out/${tag}_${t}.ncu.log 2>&1
  echo "$t rc=$?"
done
