#!/bin/bash
# 8 GPUs: NCCL's in-switch reduction (NVLS) forced for the gradient all-reduce, against the tuner's choice (ring) of the previous
# call; one GPU of the same box as the reference point
set -x
N=${1:-8}
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
NCCL_ALGO=NVLS timeout 240 $TR --master-port 29521 bench.py --gpus $N --ddp-mode end --no-parity --no-cpu-baseline > gpurun_out/r2_scale_n${N}_end_nvls.json 2> gpurun_out/r2_scale_n${N}_end_nvls.err
python tools/show_bench.py gpurun_out/r2_scale_n${N}_end_nvls.json; tail -2 gpurun_out/r2_scale_n${N}_end_nvls.err
timeout 200 python bench.py --gpus 1 --no-parity --no-cpu-baseline > gpurun_out/r2_scale_n${N}_ref1.json 2> /dev/null; python tools/show_bench.py gpurun_out/r2_scale_n${N}_ref1.json
