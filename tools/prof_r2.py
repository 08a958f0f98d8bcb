#!/usr/bin/env python
"""One launch of each hot kernel at the bench shapes inside a cudaProfilerStart/Stop bracket (warm-ups outside), for
    ncu --set full --import-source on --clock-control none --profile-from-start off -o gpurun_out/r2_prof python tools/prof_r2.py
(MMFB_LIB selects the library build).  Read here with `ncu -i ... --page raw|source --csv`."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from mmf_b200 import functional as F, lib  # noqa: E402

dev = torch.device("cuda", 0)
B, S, H, heads, I = 166, 228, 768, 12, 3072
M = B * S
bf = torch.bfloat16
x = torch.randn(M, H, device=dev).to(bf)
xi = torch.randn(M, I, device=dev).to(bf)
w_qkv = (torch.randn(3 * H, H, device=dev) * 0.02).to(bf)
w_o = (torch.randn(H, H, device=dev) * 0.02).to(bf)
w_1 = (torch.randn(I, H, device=dev) * 0.02).to(bf)
w_2 = (torch.randn(H, I, device=dev) * 0.02).to(bf)
b3, b1, bh = (torch.zeros(n, device=dev, dtype=bf) for n in (3 * H, I, H))
bits_h = F.dropout_bits((M,), H, 0.1, 1, 0, dev)
qkv = torch.randn(M, 3 * H, device=dev).to(bf)
q, k, v = qkv[:, :H], qkv[:, H:2 * H], qkv[:, 2 * H:]
lens = torch.randint(S // 2, S + 1, (B,), device=dev)          # ragged key padding, as in tools/kbench.py
mask = ((torch.arange(S, device=dev)[None] >= lens[:, None]).float() * -10000.0).contiguous()
bits_a = F.dropout_bits((B, heads, S), S, 0.1, 7, 0, dev)
dctx = torch.randn(M, H, device=dev).to(bf)
ctx, lse2, c32 = F.attention_fwd(q, k, v, B, heads, S, S, mask, bits_a, 1 / 0.9, save_lo=True)
g = torch.ones(H, device=dev, dtype=bf)
_, mean, rstd = F.layernorm_fwd(x, g, torch.zeros_like(g))
dg, db_, dbias = (torch.zeros(H, device=dev) for _ in range(3))
dq = torch.empty_like(qkv)


def attn_bwd(env):
    for k_, v_ in env.items():
        os.environ[k_] = v_
    F.attention_bwd(dctx, q, k, v, ctx, lse2, B, heads, S, S, mask, bits_a, 1 / 0.9, dq=dq[:, :H], dk=dq[:, H:2 * H],
                    dv=dq[:, 2 * H:], ctx_lo=c32)
    for k_ in env:
        os.environ.pop(k_, None)


def ln_bwd(variant):
    if variant:
        os.environ["MMFB_LN_BWD"] = variant
    F.layernorm_bwd(x, x, mean, rstd, g, dg, db_, dbias=dbias, drop_mask=bits_h, drop_scale=1 / 0.9)
    os.environ.pop("MMFB_LN_BWD", None)


def attn_fwd(env):
    for k_, v_ in env.items():
        os.environ[k_] = v_
    F.attention_fwd(q, k, v, B, heads, S, S, mask, bits_a, 1 / 0.9, save_lo=True)
    for k_ in env:
        os.environ.pop(k_, None)


work = [
    ("gemm", lambda: F.gemm(x, w_qkv, epi=lib.EPI_BIAS, bias=b3)),
    ("gemm", lambda: F.gemm(x, w_o, epi=lib.EPI_BIAS_DROP_RESID, bias=bh, aux=x, drop_mask=bits_h, drop_scale=1 / 0.9)),
    ("gemm", lambda: F.gemm(x, w_1, epi=lib.EPI_BIAS_GELU, bias=b1)),
    ("gemm", lambda: F.gemm(x, w_2, b_mn=True, epi=lib.EPI_GELU_BWD, aux=xi)),
    ("gemm", lambda: F.gemm(xi, w_2, epi=lib.EPI_BIAS_DROP_RESID, bias=bh, aux=x, drop_mask=bits_h, drop_scale=1 / 0.9)),
    ("attn", lambda: attn_fwd({})),                          # default: paired tiles, persistent
    ("attn", lambda: attn_fwd({"MMFB_ATTN_FWD": "1"})),      # one CTA per tile
    ("attn", lambda: attn_bwd({})),                          # default: persistent (+ the delta kernel)
    ("attn", lambda: attn_bwd({"MMFB_ATTN_BWD": "16"})),     # one CTA per (batch, head)
    ("ln", lambda: ln_bwd(None)),                            # default: streaming single pass
    ("ln", lambda: ln_bwd("lean")),
    ("rows", lambda: F.dropout_bits((B, heads, S), S, 0.1, 7, 0, dev)),
]
if os.environ.get("PROF_ONLY"):          # e.g. PROF_ONLY=attn,ln
    keep = os.environ["PROF_ONLY"].split(",")
    work = [w for w in work if w[0] in keep]
work = [w[1] for w in work]
for fn in work:
    fn()
torch.cuda.synchronize()
torch.cuda.profiler.start()
for fn in work:
    fn()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("profiled %d launches groups" % len(work))
