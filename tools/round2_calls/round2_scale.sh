#!/bin/bash
# multi-GPU (one box): gradient equality, then the bench in both gradient-exchange modes.   usage: round2_scale.sh N
set -x
N=${1:-2}
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
MMFB_DDP_MODE=end timeout 300 $TR --master-port 29511 tools/check_ddp.py > gpurun_out/r2_check_ddp_n${N}_end.log 2>&1; tail -3 gpurun_out/r2_check_ddp_n${N}_end.log
MMFB_DDP_MODE=bucket timeout 300 $TR --master-port 29512 tools/check_ddp.py > gpurun_out/r2_check_ddp_n${N}_bucket.log 2>&1; tail -3 gpurun_out/r2_check_ddp_n${N}_bucket.log
timeout 400 python bench.py --gpus 1 --no-parity > gpurun_out/r2_scale_n1.json 2> gpurun_out/r2_scale_n1.err; python tools/show_bench.py gpurun_out/r2_scale_n1.json
NCCL_DEBUG=INFO NCCL_DEBUG_SUBSYS=INIT,COLL,TUNING NCCL_DEBUG_FILE=gpurun_out/r2_scale_n${N}_nccl.%h.%p.log timeout 400 $TR --master-port 29513 bench.py --gpus $N --ddp-mode end --no-parity > gpurun_out/r2_scale_n${N}_end.json 2> gpurun_out/r2_scale_n${N}_end.err
python tools/show_bench.py gpurun_out/r2_scale_n${N}_end.json
cat gpurun_out/r2_scale_n${N}_nccl.*.log | grep -E "AllReduce: [0-9]{6,}|NVLS|Connected all" | cut -d" " -f3- | sort | uniq -c | sort -rn | head -12 > gpurun_out/r2_scale_n${N}_nccl.txt; rm -f gpurun_out/r2_scale_n${N}_nccl.*.log; tail -3 gpurun_out/r2_scale_n${N}_end.err
timeout 400 $TR --master-port 29514 bench.py --gpus $N --ddp-mode bucket --no-parity > gpurun_out/r2_scale_n${N}_bucket.json 2> gpurun_out/r2_scale_n${N}_bucket.err
python tools/show_bench.py gpurun_out/r2_scale_n${N}_bucket.json
timeout 400 $TR --master-port 29515 bench.py --gpus $N --ddp-mode end --ddp-payload bf16 --no-parity > gpurun_out/r2_scale_n${N}_end_bf16.json 2> gpurun_out/r2_scale_n${N}_end_bf16.err
python tools/show_bench.py gpurun_out/r2_scale_n${N}_end_bf16.json
