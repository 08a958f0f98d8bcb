#!/bin/bash
# last GPU minutes of the round: launch list of one step with the final defaults, then the bench line
set -x
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 150 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
  --log-file gpurun_out/r2_launches_step_final.csv python bench.py --profile --steps 1 --warmup 3 --no-parity 2>&1 | tail -1
python tools/agg_launches.py gpurun_out/r2_launches_step_final.csv 2>&1 | head -24
timeout 150 python bench.py --optimizer --no-parity > gpurun_out/r2_bench_final.json 2> gpurun_out/r2_bench_final.err; python tools/show_bench.py gpurun_out/r2_bench_final.json
