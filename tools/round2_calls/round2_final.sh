#!/bin/bash
# end-of-round evidence on one GPU: full GPU suite, bench lines of every workload, launch list + ncu --set full of the top kernels
set -x
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -rfE > gpurun_out/r2_gpu_tests_final.log 2>&1; tail -5 gpurun_out/r2_gpu_tests_final.log
timeout 400 python bench.py --optimizer > gpurun_out/r2_bench_final.json 2> gpurun_out/r2_bench_final.err; python tools/show_bench.py gpurun_out/r2_bench_final.json
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench_reference_arm.json 2>&1; tail -c 600 gpurun_out/r2_bench_reference_arm.json
for wl in vilbert mmbt mmft uniter_large; do
  timeout 400 python bench.py --workload $wl > gpurun_out/r2_bench_${wl}.json 2> gpurun_out/r2_bench_${wl}.err; python tools/show_bench.py gpurun_out/r2_bench_${wl}.json
done
timeout 300 python bench.py --workload mmbt --graph > gpurun_out/r2_bench_mmbt_graph.json 2> gpurun_out/r2_bench_mmbt_graph.err; python tools/show_bench.py gpurun_out/r2_bench_mmbt_graph.json
timeout 300 python tools/kbench.py --json gpurun_out/r2_kbench_final.json > gpurun_out/r2_kbench_final.log 2>&1; cat gpurun_out/r2_kbench_final.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
  --log-file gpurun_out/r2_launches_step_final.csv python bench.py --profile --steps 1 --warmup 3 --no-parity 2>&1 | tail -1
python tools/agg_launches.py gpurun_out/r2_launches_step_final.csv 2>&1 | head -30
timeout 900 ncu --set full --import-source on --clock-control none --profile-from-start off -f -o gpurun_out/r2_prof_final python tools/prof_r2.py 2>&1 | tail -2
MMFB_LIB=mmf_b200/csrc/libmmfb200_trace.so timeout 200 python tools/trace_attn.py fwd > gpurun_out/r2_trace_fwd_final.txt 2>&1
MMFB_LIB=mmf_b200/csrc/libmmfb200_trace.so timeout 200 python tools/trace_attn.py bwd > gpurun_out/r2_trace_bwd_final.txt 2>&1
timeout 300 python tools/step_gaps.py > gpurun_out/r2_step_gaps_final.txt 2>&1; head -30 gpurun_out/r2_step_gaps_final.txt
