"""TEST DOUBLE - never imported by the product.  A CPU restatement of the *contract* of each mmf_b200.functional entry
point (operand layouts, epilogues, what is accumulated where), in plain torch.  It lets the `-m "not gpu"` suite run
the engine's ORCHESTRATION (mmf_b200.engine: which kernel gets which operand, what is saved for the backward, where the
gradients are accumulated) against the oracle without a GPU.  The kernels themselves are tested on the B200 only."""
import math

import torch

from mmf_b200 import lib

LN_EPS = 1e-12
BF = torch.bfloat16
LOG2E = 1.4426950408889634


def _gelu(u):
    return u * 0.5 * (1.0 + torch.erf(u / math.sqrt(2.0)))


def _gelu_grad(u):
    return 0.5 * (1.0 + torch.erf(u / math.sqrt(2.0))) + u * torch.exp(-0.5 * u * u) / math.sqrt(2.0 * math.pi)


def unpack_keep_bits(words, n):
    w = words.to(torch.int64) & 0xFFFFFFFF
    bits = (w.unsqueeze(-1) >> torch.arange(32, dtype=torch.int64)) & 1
    return bits.reshape(*words.shape[:-1], -1)[..., :n].bool()


def pack_keep_bits(keep):
    n = keep.shape[-1]
    pad = (-n) % 32
    if pad:
        keep = torch.nn.functional.pad(keep, (0, pad))
    k = keep.reshape(*keep.shape[:-1], -1, 32).to(torch.int64)
    words = (k * (1 << torch.arange(32, dtype=torch.int64))).sum(-1)
    words = torch.where(words >= 2 ** 31, words - 2 ** 32, words)
    return words.to(torch.int32).contiguous()


def gemm(a, b, *, a_mn=False, b_mn=False, epi=lib.EPI_BIAS, bias=None, aux=None, drop_mask=None, drop_scale=1.0,
         out=None, out2=None, splits=1, block_n=0, cluster=0):
    assert a.dtype == BF and b.dtype == BF
    A = a.float().t() if a_mn else a.float()          # [M, K]
    Bm = b.float() if b_mn else b.float().t()          # [K, N]
    acc = A @ Bm
    M, N = acc.shape
    if epi == lib.EPI_ATOMIC_F32:
        if out is None:
            out = torch.zeros(M, N)
        assert out.dtype == torch.float32 and tuple(out.shape) == (M, N)
        out += acc
        return out
    if epi in (lib.EPI_BIAS, lib.EPI_BIAS_GELU, lib.EPI_BIAS_DROP_RESID, lib.EPI_BIAS_RELU) and bias is not None:
        acc = acc + bias.float()
    res2 = None
    if epi == lib.EPI_BIAS_GELU:
        res2 = _gelu(acc).to(BF)
    elif epi == lib.EPI_BIAS_RELU:
        acc = acc.clamp_min(0.0)
    elif epi == lib.EPI_BIAS_DROP_RESID:
        if drop_mask is not None:
            acc = torch.where(unpack_keep_bits(drop_mask, N), acc * drop_scale, torch.zeros_like(acc))
        if aux is not None:
            acc = acc + aux.float()
    elif epi == lib.EPI_GELU_BWD:
        acc = acc * _gelu_grad(aux.float())
    elif epi == lib.EPI_ADD_AUX:
        if aux is not None:
            acc = acc + aux.float()
    res = acc.to(BF)
    if out is not None:
        out.copy_(res)
        res = out
    return (res, res2) if epi == lib.EPI_BIAS_GELU else res


def _heads(t, B, S, h):
    return t.float().reshape(B, S, h, -1).permute(0, 2, 1, 3)          # [B, h, S, d]


def _probs(q, k, B, heads, Sq, Skv, mask):
    d = q.shape[-1] // heads
    s = _heads(q, B, Sq, heads) @ _heads(k, B, Skv, heads).transpose(-1, -2) / math.sqrt(d)
    if mask is not None:
        s = s + mask.float()[:, None, None, :]
    return s


def attention_fwd(q, k, v, B, heads, Sq, Skv, mask=None, drop_mask=None, drop_scale=1.0, out=None, save_lo=False):
    s = _probs(q, k, B, heads, Sq, Skv, mask)
    lse2 = torch.logsumexp(s, dim=-1) * LOG2E
    p = torch.softmax(s, dim=-1)
    if drop_mask is not None:
        p = torch.where(unpack_keep_bits(drop_mask, Skv), p * drop_scale, torch.zeros_like(p))
    ctx32 = (p @ _heads(v, B, Skv, heads)).permute(0, 2, 1, 3).reshape(B * Sq, -1).contiguous()
    ctx = ctx32.to(BF)
    if out is not None:
        out.copy_(ctx)
        ctx = out
    return (ctx, lse2, (ctx32 - ctx.float()).to(BF)) if save_lo else (ctx, lse2)


def attention_bwd(dctx, q, k, v, ctx, lse2, B, heads, Sq, Skv, mask=None, drop_mask=None, drop_scale=1.0, dq=None,
                  dk=None, dv=None, ctx_lo=None):
    d = q.shape[-1] // heads
    s = _probs(q, k, B, heads, Sq, Skv, mask)
    p = torch.exp2(s * LOG2E - lse2.unsqueeze(-1))              # from the saved row statistics, as the kernel does
    dO = _heads(dctx, B, Sq, heads)
    O = _heads(ctx.float() + ctx_lo.float() if ctx_lo is not None else ctx, B, Sq, heads)
    delta = (dO * O).sum(-1, keepdim=True)
    dP = dO @ _heads(v, B, Skv, heads).transpose(-1, -2)
    pd = p
    if drop_mask is not None:
        keep = unpack_keep_bits(drop_mask, Skv)
        dP = torch.where(keep, dP * drop_scale, torch.zeros_like(dP))
        pd = torch.where(keep, p * drop_scale, torch.zeros_like(p))
    dS = p * (dP - delta) / math.sqrt(d)
    merge = lambda t, S: t.permute(0, 2, 1, 3).reshape(B * S, -1).to(BF)
    rq, rk, rv = merge(dS @ _heads(k, B, Skv, heads), Sq), merge(dS.transpose(-1, -2) @ _heads(q, B, Sq, heads), Skv), \
        merge(pd.transpose(-1, -2) @ dO, Skv)
    for dst, src in ((dq, rq), (dk, rk), (dv, rv)):
        if dst is not None:
            dst.copy_(src)
    return (dq if dq is not None else rq), (dk if dk is not None else rk), (dv if dv is not None else rv)


def layernorm_fwd(y, gamma, beta, eps=LN_EPS, drop_mask=None, drop_scale=1.0, out=None):
    yf = y.float()
    mean = yf.mean(-1)
    rstd = torch.rsqrt(yf.var(-1, unbiased=False) + eps)
    x = (yf - mean[:, None]) * rstd[:, None] * gamma.float() + beta.float()
    if drop_mask is not None:
        x = torch.where(unpack_keep_bits(drop_mask, y.shape[1]), x * drop_scale, torch.zeros_like(x))
    return x.to(BF), mean, rstd


def layernorm_bwd(dx, y, mean, rstd, gamma, dgamma, dbeta, dbias=None, dx2=None, drop_mask=None, drop_scale=1.0,
                  need_dz=True):
    d = dx.float() + (dx2.float() if dx2 is not None else 0.0)
    xh = (y.float() - mean[:, None]) * rstd[:, None]
    dg = d * gamma.float()
    dy = rstd[:, None] * (dg - dg.mean(-1, keepdim=True) - xh * (dg * xh).mean(-1, keepdim=True))
    dz = dy
    if drop_mask is not None:
        dz = torch.where(unpack_keep_bits(drop_mask, y.shape[1]), dy * drop_scale, torch.zeros_like(dy))
    if dgamma is not None:
        dgamma += (d * xh).sum(0)
    if dbeta is not None:
        dbeta += d.sum(0)
    if dbias is not None:
        dbias += dz.to(BF).float().sum(0)
    dyb = dy.to(BF)
    return dyb, (dz.to(BF) if drop_mask is not None else dyb)


def colsum(x, out):
    out += x.float().sum(0)
    return out


def dropout_bits(shape_rows, ncols, p, seed, offset, device):
    g = torch.Generator().manual_seed((int(seed) * 1000003 + int(offset)) % (2 ** 63))
    keep = torch.rand(*shape_rows, ncols, generator=g) >= p
    return pack_keep_bits(keep)


def cast_f32_bf16(src, dst):
    dst.copy_(src.to(BF))


def embed_compose(M, H, srcs=(), tabs=(), device=None):
    y = torch.zeros(M, H)
    for t, rows in srcs:
        r = rows.long()
        ok = r >= 0
        y[ok] += t.float()[r[ok]]
    for t, idx in tabs:
        i = idx.long()
        ok = i >= 0
        y[ok] += t.float()[i[ok]]
    return y.to(BF)


def embed_scatter(dy, dsrcs=(), dtabs=()):
    for t, rows in dsrcs:
        r = rows.long()
        ok = r >= 0
        t[r[ok]] = dy[ok]
    for t, idx in dtabs:
        i = idx.long()
        ok = i >= 0
        t.index_add_(0, i[ok], dy.float()[ok])


def sort_indices(idx):
    s, o = torch.sort(idx.to(torch.int64), stable=True)
    return s.to(torch.int32).contiguous(), o.to(torch.int32).contiguous()


def embed_scatter_sorted(dy, dtab, sorted_idx, order):
    i = sorted_idx.long()
    ok = i >= 0
    dtab.index_add_(0, i[ok], dy.float()[order.long()[ok]])


def ce_rows(logits, labels, ignore_index, grad_scale, loss_sum, row_loss=None):
    z = logits.float()
    active = (labels != ignore_index) & (labels >= 0) & (labels < z.shape[1])
    lse = torch.logsumexp(z, dim=1)
    safe = labels.clamp(0, z.shape[1] - 1)
    li = torch.where(active, lse - z.gather(1, safe[:, None])[:, 0], torch.zeros_like(lse))
    d = torch.softmax(z, dim=1)
    d[torch.arange(z.shape[0]), safe] -= 1.0
    d = torch.where(active[:, None], d * grad_scale, torch.zeros_like(d))
    logits.copy_(d.to(BF))
    loss_sum.add_(li.sum())
    if row_loss is not None:
        row_loss.copy_(li)
    return logits


def add(a, b):
    return (a.float() + b.float()).to(BF)


def dropout_apply(x, drop_mask, drop_scale):
    keep = unpack_keep_bits(drop_mask, x.shape[-1])
    return torch.where(keep, x.float() * drop_scale, torch.zeros_like(x, dtype=torch.float32)).to(BF)


def relu_bwd(dy, y):
    return torch.where(y > 0, dy, torch.zeros_like(dy))


def adamw(param, grad, exp_avg, exp_avg_sq, groups, *, beta1, beta2, eps, mode, grad_scale=1.0, group_of_block=None,
          param_bf16=None):
    """contract of mmfb_adamw: one pass over flat fp32 buffers, hyper-parameters looked up per 8-element block"""
    assert param.numel() % 8 == 0
    gid = group_of_block.long().repeat_interleave(8) if group_of_block is not None else torch.zeros(param.numel(), dtype=torch.long)
    tab = lambda k: torch.tensor([g[k] for g in groups], dtype=torch.float32)[gid]
    lr, wd, step, bc2s = tab("lr"), tab("weight_decay"), tab("step_size"), tab("bc2_sqrt")
    g = grad * grad_scale
    if mode == 0:
        exp_avg.mul_(beta1).add_(g, alpha=1.0 - beta1)
        exp_avg_sq.mul_(beta2).addcmul_(g, g, value=1.0 - beta2)
        param.sub_(step * (exp_avg / (exp_avg_sq.sqrt() + eps)))
        param.sub_(torch.where(wd > 0, lr * wd * param, torch.zeros_like(param)))
    else:
        param.mul_(1.0 - lr * wd)
        exp_avg.lerp_(g, 1.0 - beta1)
        exp_avg_sq.mul_(beta2).addcmul_(g, g, value=1.0 - beta2)
        param.sub_(step * (exp_avg / (exp_avg_sq.sqrt() / bc2s + eps)))
    if param_bf16 is not None:
        param_bf16.copy_(param.to(BF))


def gelu_bwd(dh, u):
    return (dh.float() * _gelu_grad(u.float())).to(BF)
