#!/bin/bash
# 8 GPUs: NCCL's in-switch reduction (NVLS) for the all-reduce only (per-function NCCL_ALGO syntax)
set -x
N=${1:-8}
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
NCCL_ALGO="allreduce:nvls" NCCL_DEBUG=INFO NCCL_DEBUG_SUBSYS=COLL,TUNING NCCL_DEBUG_FILE=gpurun_out/r2_scale_n${N}_nvls_nccl.%h.%p.log timeout 200 $TR --master-port 29531 bench.py --gpus $N --ddp-mode end --no-parity --no-cpu-baseline --steps 16 --warmup 4 > gpurun_out/r2_scale_n${N}_end_nvls.json 2> gpurun_out/r2_scale_n${N}_end_nvls.err
python tools/show_bench.py gpurun_out/r2_scale_n${N}_end_nvls.json; grep -E "Error|error" gpurun_out/r2_scale_n${N}_end_nvls.err | head -3
cat gpurun_out/r2_scale_n${N}_nvls_nccl.*.log | grep -E "AllReduce: [0-9]{6,}" | cut -d" " -f3- | sort | uniq -c | sort -rn | head -4 > gpurun_out/r2_scale_n${N}_nvls_nccl.txt; rm -f gpurun_out/r2_scale_n${N}_nvls_nccl.*.log; cat gpurun_out/r2_scale_n${N}_nvls_nccl.txt
