#!/bin/bash
# whole-step CUDA graph: parity test, then the small configurations eager vs graph
set -x
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_next_gpu.py tests/test_rowops_gpu.py -m gpu -q -x -rfE > gpurun_out/r2_graph_tests13.log 2>&1; tail -6 gpurun_out/r2_graph_tests13.log
for wl in mmbt vilbert; do
  timeout 400 python bench.py --workload $wl --no-cpu-baseline > gpurun_out/r2_bench_${wl}_eager.json 2> gpurun_out/r2_bench_${wl}_eager.err; python tools/show_bench.py gpurun_out/r2_bench_${wl}_eager.json; tail -2 gpurun_out/r2_bench_${wl}_eager.err
  timeout 400 python bench.py --workload $wl --graph --no-cpu-baseline > gpurun_out/r2_bench_${wl}_graph.json 2> gpurun_out/r2_bench_${wl}_graph.err; python tools/show_bench.py gpurun_out/r2_bench_${wl}_graph.json; tail -2 gpurun_out/r2_bench_${wl}_graph.err
done
timeout 400 python bench.py --graph --no-cpu-baseline > gpurun_out/r2_bench_visual_bert_graph.json 2> gpurun_out/r2_bench_visual_bert_graph.err; python tools/show_bench.py gpurun_out/r2_bench_visual_bert_graph.json; tail -2 gpurun_out/r2_bench_visual_bert_graph.err
