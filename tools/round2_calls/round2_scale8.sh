#!/bin/bash
# 8 GPUs of one box (charged 8x): gradient equality once, then the bench in the default exchange mode (+ the bf16 payload)
set -x
N=${1:-8}
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
MMFB_DDP_MODE=end timeout 150 $TR --master-port 29511 tools/check_ddp.py > gpurun_out/r2_check_ddp_n${N}_end.log 2>&1; tail -2 gpurun_out/r2_check_ddp_n${N}_end.log
NCCL_DEBUG=INFO NCCL_DEBUG_SUBSYS=INIT,COLL,TUNING NCCL_DEBUG_FILE=gpurun_out/r2_scale_n${N}_nccl.%h.%p.log timeout 240 $TR --master-port 29513 bench.py --gpus $N --ddp-mode end --no-parity --no-cpu-baseline > gpurun_out/r2_scale_n${N}_end.json 2> gpurun_out/r2_scale_n${N}_end.err
python tools/show_bench.py gpurun_out/r2_scale_n${N}_end.json; tail -2 gpurun_out/r2_scale_n${N}_end.err
cat gpurun_out/r2_scale_n${N}_nccl.*.log | grep -E "AllReduce: [0-9]{6,}|NVLS|Connected all" | cut -d" " -f3- | sort | uniq -c | sort -rn | head -12 > gpurun_out/r2_scale_n${N}_nccl.txt; rm -f gpurun_out/r2_scale_n${N}_nccl.*.log
timeout 240 $TR --master-port 29515 bench.py --gpus $N --ddp-mode end --ddp-payload bf16 --no-parity --no-cpu-baseline > gpurun_out/r2_scale_n${N}_end_bf16.json 2> gpurun_out/r2_scale_n${N}_end_bf16.err
python tools/show_bench.py gpurun_out/r2_scale_n${N}_end_bf16.json
