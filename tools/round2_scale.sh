#!/bin/bash
# multi-GPU (one box): gradient equality, then the bench in both gradient-exchange modes.   usage: round2_scale.sh N
set -x
N=${1:-2}
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
MMFB_DDP_MODE=end timeout 300 $TR --master-port 29511 tools/check_ddp.py > gpurun_out/r2_check_ddp_n${N}_end.log 2>&1; tail -3 gpurun_out/r2_check_ddp_n${N}_end.log
MMFB_DDP_MODE=bucket timeout 300 $TR --master-port 29512 tools/check_ddp.py > gpurun_out/r2_check_ddp_n${N}_bucket.log 2>&1; tail -3 gpurun_out/r2_check_ddp_n${N}_bucket.log
timeout 400 python bench.py --gpus 1 --no-parity > gpurun_out/r2_scale_n1.json 2> gpurun_out/r2_scale_n1.err; python tools/show_bench.py gpurun_out/r2_scale_n1.json
NCCL_DEBUG=INFO NCCL_DEBUG_SUBSYS=INIT,COLL,TUNING timeout 400 $TR --master-port 29513 bench.py --gpus $N --ddp-mode end --no-parity > gpurun_out/r2_scale_n${N}_end.json 2> gpurun_out/r2_scale_n${N}_end.err
python tools/show_bench.py gpurun_out/r2_scale_n${N}_end.json
grep -E "NVLS|Channel|Algo|Using|nChannels" gpurun_out/r2_scale_n${N}_end.err | sort | uniq -c | sort -rn | head -12 > gpurun_out/r2_scale_n${N}_nccl.txt; grep -v "NCCL INFO" gpurun_out/r2_scale_n${N}_end.err | tail -5
timeout 400 $TR --master-port 29514 bench.py --gpus $N --ddp-mode bucket --no-parity > gpurun_out/r2_scale_n${N}_bucket.json 2> gpurun_out/r2_scale_n${N}_bucket.err
python tools/show_bench.py gpurun_out/r2_scale_n${N}_bucket.json
timeout 400 $TR --master-port 29515 bench.py --gpus $N --ddp-mode end --no-parity > gpurun_out/r2_scale_n${N}_end2.json 2> gpurun_out/r2_scale_n${N}_end2.err
python tools/show_bench.py gpurun_out/r2_scale_n${N}_end2.json
