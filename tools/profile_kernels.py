"""Runs the dominant kernels alone (for `ncu --set full -k regex:...`): FFN-up GEMM (+bias+GELU), GELU' dgrad GEMM,
attention fwd / bwd at the config-2 shapes (B=64, S=228, H=768)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mmf_b200 import functional as F, lib

B, S, H, I, heads = 64, 228, 768, 3072, 12
M = B * S
dev = "cuda"
a = torch.randn(M, H, device=dev).to(torch.bfloat16)
w1 = (torch.randn(I, H, device=dev) * 0.02).to(torch.bfloat16)
b1 = torch.zeros(I, device=dev, dtype=torch.bfloat16)
for _ in range(3):
    u, h = F.gemm(a, w1, epi=lib.EPI_BIAS_GELU, bias=b1)
    qkv = F.gemm(a, w1[:2304], epi=lib.EPI_BIAS, bias=b1[:2304])
    dz = torch.randn(M, H, device=dev).to(torch.bfloat16)
    w2 = (torch.randn(H, I, device=dev) * 0.02).to(torch.bfloat16)
    du = F.gemm(dz, w2, b_mn=True, epi=lib.EPI_GELU_BWD, aux=u)
    mask = torch.zeros(B, S, device=dev)
    bits = F.dropout_bits((B, heads, S), S, 0.1, 1, 0, dev)
    ctx, lse = F.attention_fwd(qkv[:, :H], qkv[:, H:2 * H], qkv[:, 2 * H:], B, heads, S, S, mask, bits, 1 / 0.9)
    dq, dk, dv = F.attention_bwd(dz, qkv[:, :H], qkv[:, H:2 * H], qkv[:, 2 * H:], ctx, lse, B, heads, S, S, mask, bits, 1 / 0.9)
torch.cuda.synchronize()
print("done")
