#!/bin/bash
# First GPU call of round 2 (one gpurun): the WHOLE -m gpu suite with the staged tests enabled (no -x: every failure is
# listed), parity of every staged kernel variant, then one sweep that times the product library once and each variant
# against it on the same box, then the ViLBERT throughput script.
#     python tools/ab.py build x2 -DMMFB_F32X2=1        (here, before the call)
#     tools/gpurun_retry.sh 1500 'bash tools/round2_first_call.sh > gpurun_out/round2_first_call.log 2>&1'
set -x
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv
timeout 900 python -m pytest tests -m gpu -q -rfEP > gpurun_out/r2_gpu_tests_full.log 2>&1; tail -40 gpurun_out/r2_gpu_tests_full.log
# staged tests one function per process: a trapped kernel (mbarrier watchdog) must not poison the tests after it
for t in $(grep -o "^def test_[a-z0-9_]*" tests/test_staged_gpu.py | sed 's/def //'); do
  MMFB_STAGED_TESTS=1 timeout 200 python -m pytest tests/test_staged_gpu.py -m gpu -q -k "$t" 2>&1 | tail -4
done
MMFB_LIB=$PWD/mmf_b200/csrc/libmmfb200_x2.so timeout 300 python -m pytest tests -m gpu -q -x 2>&1 | tail -5
MMFB_LIB=$PWD/mmf_b200/csrc/libmmfb200_pf.so timeout 300 python -m pytest tests/test_gemm_gpu.py tests/test_encoder_gpu.py -m gpu -q -x 2>&1 | tail -5
MMFB_LN_BWD=lean timeout 300 python -m pytest tests/test_rowops_gpu.py tests/test_encoder_gpu.py -m gpu -q -x 2>&1 | tail -5
MMFB_LN_BWD=tile timeout 300 python -m pytest tests/test_rowops_gpu.py tests/test_encoder_gpu.py tests/test_frontends_gpu.py -m gpu -q -x 2>&1 | tail -5
MMFB_ATTN_FWD=2 timeout 300 python -m pytest tests/test_attention_gpu.py tests/test_encoder_gpu.py -m gpu -q -x 2>&1 | tail -5
MMFB_ATTN_BWD_OVERLAP=1 timeout 300 python -m pytest tests/test_attention_gpu.py tests/test_encoder_gpu.py -m gpu -q -x 2>&1 | tail -5
MMFB_ATTN_BWD=16 timeout 300 python -m pytest tests/test_attention_gpu.py tests/test_encoder_gpu.py -m gpu -q -x 2>&1 | tail -5
MMFB_DROPOUT_ASYNC=1 MMFB_SIDE_REDUCE=1 timeout 300 python -m pytest tests/test_encoder_gpu.py tests/test_visual_bert_gpu.py -m gpu -q -x 2>&1 | tail -5
timeout 900 python tools/ab.py sweep x2 pf pfx2 lnlean:MMFB_LN_BWD=lean lntile:MMFB_LN_BWD=tile rng:MMFB_DROPOUT_ASYNC=1 side:MMFB_SIDE_REDUCE=1 \
    attn2:MMFB_ATTN_FWD=2 bwdovl:MMFB_ATTN_BWD_OVERLAP=1 bwd16:MMFB_ATTN_BWD=16 \
    all:MMFB_LN_BWD=tile,MMFB_DROPOUT_ASYNC=1,MMFB_SIDE_REDUCE=1,MMFB_ATTN_FWD=2,MMFB_ATTN_BWD=16 \
    allpfx2:MMFB_LN_BWD=tile,MMFB_DROPOUT_ASYNC=1,MMFB_SIDE_REDUCE=1,MMFB_ATTN_FWD=2,MMFB_ATTN_BWD=16 --steps 12
for w in vilbert mmbt mmft uniter_large; do
  timeout 400 python bench.py --workload $w --steps 8 --warmup 3 > gpurun_out/r2_bench_$w.json 2> gpurun_out/r2_bench_$w.err
  tail -c 1500 gpurun_out/r2_bench_$w.json; tail -3 gpurun_out/r2_bench_$w.err
done
timeout 400 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_visual_bert_first.json 2> gpurun_out/r2_bench_vb.err; tail -c 2500 gpurun_out/r2_bench_visual_bert_first.json
