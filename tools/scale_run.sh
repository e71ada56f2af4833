#!/bin/bash
# One multi-GPU measurement session (run under `gpurun --gpus N`): the multi-GPU tests, then bench.py for the three workloads
# (default weak scaling at the driver's 20 steps and at 500; BASELINE configs[3] and configs[4] at full size, strong scaling).
#   bash tools/scale_run.sh <N> <tag>
N=$1; tag=$2
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29577"
[ "$N" = 1 ] && TR="python"
python -m pytest tests/test_multi_gpu.py -m gpu -x -q > gpurun_out/${tag}_pytest_${N}gpu.log 2>&1; tail -2 gpurun_out/${tag}_pytest_${N}gpu.log
run() { name=$1; shift; $TR bench.py --gpus $N "$@" > gpurun_out/${tag}_${name}_${N}gpu.json 2> gpurun_out/${tag}_${name}_${N}gpu.err || tail -5 gpurun_out/${tag}_${name}_${N}gpu.err; python tools/benchq.py ${name}_${N}gpu < gpurun_out/${tag}_${name}_${N}gpu.json; }
run default20 --steps 20 --warmup 3
run default500 --steps 500 --warmup 3
run sdxl --shape sdxl --total-latents 65536 --steps 5
run plk --per-latent-keys --total-latents 1048576 --steps 5
