"""Thin Python wrappers over the C ABI: one function per exported kernel entry point.

Each wrapper validates dtypes/devices, allocates outputs with torch (so the caching allocator and
autograd own every buffer) and enqueues the kernel on the current CUDA stream.  Nothing here
computes on the CPU.
"""
import ctypes

import torch

from . import lib
from .lib import LIB, GemmArgs, _ptr, _stream_ptr, check


def _req(t, dtype, name):
    if t is None:
        return
    if not t.is_cuda:
        raise RuntimeError("%s must be a CUDA tensor: the B200 path has no CPU fallback" % name)
    if t.dtype != dtype:
        raise ValueError("%s must be %s, got %s" % (name, dtype, t.dtype))
    if t.stride(-1) != 1:
        raise ValueError("%s must be contiguous in its last dimension" % name)


def gemm(a, b, *, a_mn=False, b_mn=False, epi=lib.EPI_BIAS, bias=None, aux=None, drop_mask=None,
         drop_scale=1.0, out=None, out2=None, splits=1, block_n=0):
    """C[M,N] = epi(A[M,K] @ B[N,K]^T) on the tcgen05 GEMM.

    a: [M,K] (a_mn=False) or [K,M] (a_mn=True); b: [N,K] (b_mn=False) or [K,N] (b_mn=True); bf16.
    Returns C (and C2 for EPI_BIAS_GELU).  For EPI_ATOMIC_F32 `out` (fp32 [M,N]) is accumulated into.
    """
    _req(a, torch.bfloat16, "a")
    _req(b, torch.bfloat16, "b")
    _req(bias, torch.bfloat16, "bias")
    _req(aux, torch.bfloat16, "aux")
    if a.dim() != 2 or b.dim() != 2:
        raise ValueError("gemm operands must be 2-D")
    M, K = (a.shape[1], a.shape[0]) if a_mn else (a.shape[0], a.shape[1])
    N, Kb = (b.shape[1], b.shape[0]) if b_mn else (b.shape[0], b.shape[1])
    if K != Kb:
        raise ValueError("gemm: inner dimensions differ (%d vs %d)" % (K, Kb))
    if epi == lib.EPI_ATOMIC_F32:
        if out is None:
            out = torch.zeros(M, N, dtype=torch.float32, device=a.device)
        _req(out, torch.float32, "out")
    else:
        if out is None:
            out = torch.empty(M, N, dtype=torch.bfloat16, device=a.device)
        _req(out, torch.bfloat16, "out")
        if epi == lib.EPI_BIAS_GELU and out2 is None:
            out2 = torch.empty(M, N, dtype=torch.bfloat16, device=a.device)
    args = GemmArgs()
    args.A, args.lda, args.a_mn = a.data_ptr(), a.stride(0), int(a_mn)
    args.B, args.ldb, args.b_mn = b.data_ptr(), b.stride(0), int(b_mn)
    args.C, args.ldc = out.data_ptr(), out.stride(0)
    args.C2 = out2.data_ptr() if out2 is not None else None
    args.bias = bias.data_ptr() if bias is not None else None
    args.aux, args.ldaux = (aux.data_ptr(), aux.stride(0)) if aux is not None else (None, 0)
    if drop_mask is not None:
        _req(drop_mask, torch.int32, "drop_mask")
        args.drop_mask, args.ldmask = drop_mask.data_ptr(), drop_mask.stride(0)
    args.drop_scale = float(drop_scale)
    args.M, args.N, args.K = M, N, K
    args.epi, args.splits, args.block_n = int(epi), int(splits), int(block_n)
    check(LIB.mmfb_gemm(ctypes.byref(args), _stream_ptr()))
    if epi == lib.EPI_BIAS_GELU:
        return out, out2
    return out
