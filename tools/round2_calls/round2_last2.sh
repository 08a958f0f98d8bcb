#!/bin/bash
set -x
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 170 python bench.py --optimizer > gpurun_out/r2_bench_final_graph.json 2> gpurun_out/r2_bench_final_graph.err; python tools/show_bench.py gpurun_out/r2_bench_final_graph.json; tail -3 gpurun_out/r2_bench_final_graph.err
