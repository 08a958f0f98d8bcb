"""Times every GEMM shape of one BERT layer (fwd + bwd, config 2, B=64) with block_n 128 and 256: TFLOP/s per shape."""
import sys, os, statistics
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mmf_b200 import functional as F, lib
_orig_gemm = F.gemm
def _gemm(*a, **k):
    k.setdefault('cluster', int(os.environ.get('MMFB_SWEEP_CLUSTER', '0')))
    return _orig_gemm(*a, **k)
F.gemm = _gemm
from mmf_b200.engine import best_splits

B, S, H, I = int(os.environ.get("B", 64)), 228, 768, 3072
M = B * S
dev = "cuda"
r = lambda *s: (torch.randn(*s, device=dev) * 0.05).to(torch.bfloat16)
x, w_qkv, w_o, w1, w2 = r(M, H), r(3 * H, H), r(H, H), r(I, H), r(H, I)
b3, bh, bi = r(3 * H), r(H), r(I)
u, dz, du, dqkv = r(M, I), r(M, H), r(M, I), r(M, 3 * H)
flush = torch.empty(200 << 20, dtype=torch.uint8, device=dev)
cases = [
    ("qkv   fwd bias      ", lambda bn: F.gemm(x, w_qkv, epi=lib.EPI_BIAS, bias=b3, block_n=bn), 2 * M * H * 3 * H),
    ("oproj fwd drop+resid", lambda bn: F.gemm(x, w_o, epi=lib.EPI_BIAS_DROP_RESID, bias=bh, aux=x, block_n=bn), 2 * M * H * H),
    ("ffn1  fwd bias+gelu ", lambda bn: F.gemm(x, w1, epi=lib.EPI_BIAS_GELU, bias=bi, block_n=bn), 2 * M * H * I),
    ("ffn2  fwd drop+resid", lambda bn: F.gemm(u, w2, epi=lib.EPI_BIAS_DROP_RESID, bias=bh, aux=x, block_n=bn), 2 * M * H * I),
    ("ffn2  dgrad gelu'   ", lambda bn: F.gemm(dz, w2, b_mn=True, epi=lib.EPI_GELU_BWD, aux=u, block_n=bn), 2 * M * H * I),
    ("ffn1  dgrad +resid  ", lambda bn: F.gemm(du, w1, b_mn=True, epi=lib.EPI_ADD_AUX, aux=x, block_n=bn), 2 * M * H * I),
    ("oproj dgrad         ", lambda bn: F.gemm(dz, w_o, b_mn=True, epi=lib.EPI_BIAS, block_n=bn), 2 * M * H * H),
    ("qkv   dgrad +resid  ", lambda bn: F.gemm(dqkv, w_qkv, b_mn=True, epi=lib.EPI_ADD_AUX, aux=x, block_n=bn), 2 * M * H * 3 * H),
]
wg = [("wgrad ffn2 [768,3072]", dz, u), ("wgrad ffn1 [3072,768]", du, x), ("wgrad oproj [768,768]", dz, x), ("wgrad qkv [2304,768]", dqkv, x)]


def timeit(fn, n=8):
    ts = []
    for i in range(n + 2):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        if i >= 2:
            ts.append(e0.elapsed_time(e1))
    return statistics.median(ts)


import functools
for name, fn, fl in cases:
    out = []
    for cl in (1, 2):
        os.environ["MMFB_SWEEP_CLUSTER"] = str(cl)
        ms = timeit(lambda: fn(256))
        out.append("cluster=%d %6.1f us %6.0f TF" % (cl - 1, ms * 1e3, fl / ms / 1e9))
    print(name, " | ".join(out))
for name, a, b in wg:
    Mo, No = a.shape[1], b.shape[1]
    g = torch.zeros(Mo, No, device=dev)
    fl = 2 * M * Mo * No
    out = []
    sp = best_splits(Mo, No, M)
    for cl in (1, 2):
        ms = timeit(lambda: F.gemm(a, b, a_mn=True, b_mn=True, epi=lib.EPI_ATOMIC_F32, out=g, splits=sp, cluster=cl))
        out.append("cluster=%d s%d %6.1f us %5.0f TF" % (cl - 1, sp, ms * 1e3, fl / ms / 1e9))
    print(name, " | ".join(out))
