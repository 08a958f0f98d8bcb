"""Runs only the bench's dominant kernel (FFN-up GEMM + bias + GELU, two outputs) at the bench batch, for
`ncu --set full -k regex:gemm_kernel` (DRAM traffic per launch -> bench.py roofline.traffic)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mmf_b200 import functional as F, lib
B = int(os.environ.get("MMFB_BENCH_BATCH", "166"))
M, H, I = B * 228, 768, 3072
a = torch.randn(M, H, device="cuda").to(torch.bfloat16)
w = (torch.randn(I, H, device="cuda") * 0.02).to(torch.bfloat16)
b = torch.zeros(I, device="cuda", dtype=torch.bfloat16)
o1 = torch.empty(M, I, device="cuda", dtype=torch.bfloat16); o2 = torch.empty_like(o1)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for _ in range(4):
    flush.zero_()
    F.gemm(a, w, epi=lib.EPI_BIAS_GELU, bias=b, out=o1, out2=o2)
torch.cuda.synchronize()
