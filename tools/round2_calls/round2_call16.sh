#!/bin/bash
# whole GPU suite (incl. the repeated-launch regression and the graph test), then the suite again twice for flakiness
set -x
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
for i in 1 2 3; do timeout 900 python -m pytest tests -m gpu -q -rfE > gpurun_out/r2_gpu_tests_full16_$i.log 2>&1; tail -3 gpurun_out/r2_gpu_tests_full16_$i.log; done
timeout 400 python bench.py > gpurun_out/r2_bench_call16.json 2> gpurun_out/r2_bench_call16.err; python tools/show_bench.py gpurun_out/r2_bench_call16.json
