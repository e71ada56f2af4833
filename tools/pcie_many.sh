#!/bin/bash
# PCIe ceilings with N GPUs busy at once: tools/pciebench.py on GPUs 0..N-1 concurrently.  Usage: bash tools/pcie_many.sh N tag
n=$1; tag=$2
mkdir -p gpurun_out
for g in $(seq 0 $((n-1))); do CUDA_VISIBLE_DEVICES=$g python tools/pciebench.py > gpurun_out/${tag}_n${n}_g$g.json & done
wait
cat gpurun_out/${tag}_n${n}_g*.json | python -c "
import sys, json
rows=[json.loads(l) for l in sys.stdin if l.strip()]
f=lambda k: [round(r[k],1) for r in rows]
print(json.dumps({'gpus_busy': len(rows), 'h2d_GBps': f('h2d_GBps'), 'd2h_GBps': f('d2h_GBps'), 'both_each_GBps': f('both_each_GBps'), 'both_ms': f('both_ms')}))"
