#!/usr/bin/env python
"""Phase timeline of the persistent attention kernels (development tool; needs the trace build of the library):
    python tools/ab.py build trace -DMMFB_TRACE=1        (here)
    MMFB_LIB=mmf_b200/csrc/libmmfb200_trace.so python tools/trace_attn.py [fwd|bwd]      (on the GPU box)
CTA 0 stamps clock64() at fixed points of its first items (MMFB_TR in csrc/attention.cu); this prints, per item, the cycles
between consecutive stamps of each role."""
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from mmf_b200 import functional as F, lib  # noqa: E402

which = sys.argv[1] if len(sys.argv) > 1 else "fwd"
dev = torch.device("cuda", 0)
torch.manual_seed(0)
B, S, H, heads = 166, 228, 768, 12
M = B * S
bf = torch.bfloat16
qkv = torch.randn(M, 3 * H, device=dev).to(bf)
q, k, v = qkv[:, :H], qkv[:, H:2 * H], qkv[:, 2 * H:]
lens = torch.randint(S // 2, S + 1, (B,), device=dev)
mask = ((torch.arange(S, device=dev)[None] >= lens[:, None]).float() * -10000.0).contiguous()
bits = F.dropout_bits((B, heads, S), S, 0.1, 7, 0, dev)
dctx = torch.randn(M, H, device=dev).to(bf)
ctx, lse2, c32 = F.attention_fwd(q, k, v, B, heads, S, S, mask, bits, 1 / 0.9, save_lo=True)
dq = torch.empty_like(qkv)
L = lib.LIB
L.mmfb_trace_read.argtypes = [ctypes.POINTER(ctypes.c_longlong), ctypes.c_int]
N = 4 * 64 * 16
buf = (ctypes.c_longlong * N)()


def run():
    if which == "fwd":
        F.attention_fwd(q, k, v, B, heads, S, S, mask, bits, 1 / 0.9, save_lo=True)
    else:
        F.attention_bwd(dctx, q, k, v, ctx, lse2, B, heads, S, S, mask, bits, 1 / 0.9, dq=dq[:, :H], dk=dq[:, H:2 * H],
                        dv=dq[:, 2 * H:], ctx_lo=c32)


for _ in range(3):
    run()
torch.cuda.synchronize()
L.mmfb_trace_clear()
run()
torch.cuda.synchronize()
assert L.mmfb_trace_read(buf, N) == 0
t = [[[buf[(r * 64 + n) * 16 + s] for s in range(16)] for n in range(64)] for r in range(4)]
base = min(x for r in t for n in r for x in n if x > 0)
print("kernel: attention %s, CTA 0; cycles relative to its first stamp (0 = slot not stamped)" % which)
for r in range(4):
    if not any(any(n) for n in t[r]):
        continue
    print("role %d" % r)
    for n in range(16):
        row = t[r][n]
        if not any(row):
            continue
        print("  item %2d: " % n + " ".join("%7d" % (x - base if x else 0) for x in row))
