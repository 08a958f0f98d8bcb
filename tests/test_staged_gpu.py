"""Checks for STAGED (off-by-default, not yet measured) variants.  Skipped unless MMFB_STAGED_TESTS=1, so that the default
`-m gpu` run only covers what the product library executes.  The kernel variants themselves (MMFB_LN_BWD=lean,
MMFB_ATTN_FWD=2, the -DMMFB_F32X2 build) are exercised by running the ordinary suite with the switch set
(tools/ab.py run NAME --env KEY=VALUE --tests)."""
import os

import pytest
import torch

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(os.environ.get("MMFB_STAGED_TESTS") != "1", reason="staged variants: set MMFB_STAGED_TESTS=1")]


def test_async_dropout_state_draws_the_same_bits():
    from mmf_b200 import engine as E
    sites = [((4, 12, 228), 228, 0.1), ((4 * 228,), 768, 0.1), ((4 * 228,), 768, 0.1)]
    sync = E.DropoutState(1234)
    ref = [sync.bits(r, n, p, "cuda") for r, n, p in sites + sites]
    a = E.AsyncDropoutState(1234)
    a.prefetch(sites, torch.device("cuda"))
    got = []
    for k, (r, n, p) in enumerate(sites + sites):
        if k == 1:
            a.prefetch(sites, torch.device("cuda"))     # the next layer, announced while this one is being consumed
        got.append(a.bits(r, n, p, "cuda"))
    torch.cuda.synchronize()
    for x, y in zip(ref, got):
        assert torch.equal(x, y)
    with pytest.raises(RuntimeError):
        a.prefetch(sites, torch.device("cuda"))
        a.bits((1,), 32, 0.1, "cuda")                   # asked out of order


@pytest.mark.parametrize("variant,M,H", [("lean", 1000, 768), ("tile", 1000, 768), ("tile", 37848, 768), ("tile", 333, 1024),
                                         ("tile", 50, 128), ("tile", 7, 512)])
def test_lean_layernorm_backward_matches_the_default_pair(variant, M, H):
    from mmf_b200 import functional as F
    torch.manual_seed(0)
    dx = torch.randn(M, H, device="cuda").to(torch.bfloat16)
    dx2 = torch.randn(M, H, device="cuda").to(torch.bfloat16)
    y = torch.randn(M, H, device="cuda").to(torch.bfloat16)
    g = (1.0 + 0.1 * torch.randn(H, device="cuda")).to(torch.bfloat16)
    b = torch.zeros(H, device="cuda", dtype=torch.bfloat16)
    _, mean, rstd = F.layernorm_fwd(y, g, b)
    bits = F.dropout_bits((M,), H, 0.1, 3, 0, "cuda")

    def run(variant, with_dx2, with_drop):
        if variant:
            os.environ["MMFB_LN_BWD"] = variant
        else:
            os.environ.pop("MMFB_LN_BWD", None)
        dg, db, dbias = (torch.zeros(H, device="cuda") for _ in range(3))
        dy, dz = F.layernorm_bwd(dx, y, mean, rstd, g, dg, db, dbias=dbias, dx2=dx2 if with_dx2 else None,
                                 drop_mask=bits if with_drop else None, drop_scale=1.0 / 0.9)
        torch.cuda.synchronize()
        return dy.float(), dz.float(), dg, db, dbias
    try:
        for with_dx2 in (False, True):
            for with_drop in (False, True):
                ref = run(None, with_dx2, with_drop)
                got = run(variant, with_dx2, with_drop)
                for r, t in zip(ref[:2], got[:2]):      # same row arithmetic; FMA contraction may flip a last bf16 bit
                    assert (r - t).abs().max() <= 1e-2 * r.abs().max() and (r != t).float().mean() < 0.01
                for r, t in zip(ref[2:], got[2:]):                                         # sums: different order
                    assert (r - t).abs().max() <= 1e-3 * r.abs().max().clamp_min(1.0)
    finally:
        os.environ.pop("MMFB_LN_BWD", None)


@pytest.mark.parametrize("Sq,Skv,drop", [(228, 228, True), (128, 256, False), (100, 36, True), (256, 17, False)])
def test_pipelined_attention_forward_matches_the_default_kernel(Sq, Skv, drop):
    from mmf_b200 import functional as F
    torch.manual_seed(Sq + Skv)
    B, heads, d = 5, 3, 64
    W = heads * d
    q = torch.randn(B * Sq, W, device="cuda").to(torch.bfloat16)
    k = torch.randn(B * Skv, W, device="cuda").to(torch.bfloat16)
    v = torch.randn(B * Skv, W, device="cuda").to(torch.bfloat16)
    mask = torch.zeros(B, Skv, device="cuda")
    mask[1, Skv // 2:] = -10000.0
    mask[2, :] = -10000.0                                   # fully masked sample: uniform softmax
    bits = F.dropout_bits((B, heads, Sq), Skv, 0.1, 5, 0, "cuda") if drop else None

    def run(variant):
        if variant:
            os.environ["MMFB_ATTN_FWD"] = variant
        else:
            os.environ.pop("MMFB_ATTN_FWD", None)
        ctx, lse2, c32 = F.attention_fwd(q, k, v, B, heads, Sq, Skv, mask, bits, 1.0 / 0.9 if drop else 1.0, save_fp32=True)
        torch.cuda.synchronize()
        return ctx.float(), lse2, c32
    try:
        ref, got = run(None), run("2")
    finally:
        os.environ.pop("MMFB_ATTN_FWD", None)
    assert torch.isfinite(got[0]).all()
    assert (ref[1] - got[1]).abs().max() < 1e-3             # row statistics (log2 domain)
    assert (ref[2] - got[2]).abs().max() <= 2e-2 * ref[2].abs().max()      # same P (bf16) x V, different summation order
    assert (ref[0] - got[0]).abs().max() <= 2e-2 * ref[0].abs().max()


@pytest.mark.parametrize("mode", [0, 1])
def test_adamw_kernel_matches_the_oracle(mode):
    """mmfb_adamw over a flat buffer with two hyper-parameter groups + a frozen one, 3 steps, vs the fp32 oracle"""
    import math
    from mmf_b200 import functional as F
    from oracle import fusion_oracle as O
    torch.manual_seed(mode)
    n = 8 * 1000
    p = torch.randn(n, device="cuda")
    m, v = torch.zeros_like(p), torch.zeros_like(p)
    gid = torch.randint(0, 3, (n // 8,), device="cuda", dtype=torch.uint8)
    lr, b1, b2, eps = 3e-3, 0.9, 0.98, 1e-6
    wds = [0.01, 0.0, 0.0]
    ref_p, ref_m, ref_v = p.cpu().clone(), m.cpu().clone(), v.cpu().clone()
    sel = gid.cpu().long().repeat_interleave(8)
    for step in range(1, 4):
        g = torch.randn(n, device="cuda") * 0.1
        hps = []
        for gi in range(3):
            if gi == 2:
                hps.append({"lr": 0.0, "weight_decay": 0.0, "step_size": 0.0, "bc2_sqrt": 1.0})      # frozen group
            elif mode == 0:
                hps.append({"lr": lr, "weight_decay": wds[gi], "bc2_sqrt": 1.0,
                            "step_size": lr * math.sqrt(1 - b2 ** step) / (1 - b1 ** step)})
            else:
                hps.append({"lr": lr, "weight_decay": wds[gi], "step_size": lr / (1 - b1 ** step),
                            "bc2_sqrt": math.sqrt(1 - b2 ** step)})
        F.adamw(p, g, m, v, hps, beta1=b1, beta2=b2, eps=eps, mode=mode, grad_scale=0.5, group_of_block=gid)
        gc = g.cpu() * 0.5
        for gi in range(2):
            idx = sel == gi
            pp, mm, vv = ref_p[idx], ref_m[idx], ref_v[idx]
            (O.adamw_step_transformers if mode == 0 else O.adamw_step_torch)(pp, gc[idx], mm, vv, step, lr, b1, b2, eps, wds[gi])
            ref_p[idx], ref_m[idx], ref_v[idx] = pp, mm, vv
        torch.cuda.synchronize()
        idx = sel < 2
        assert (p.cpu()[idx] - ref_p[idx]).abs().max() <= 2e-6 * ref_p[idx].abs().max()
        assert torch.equal(p.cpu()[sel == 2], ref_p[sel == 2])          # frozen blocks are bit-identical


@pytest.mark.parametrize("Sq,Skv,drop", [(228, 228, True), (256, 130, False), (100, 256, True)])
def test_overlapped_fused_attention_backward_matches_the_default(Sq, Skv, drop):
    from mmf_b200 import functional as F
    torch.manual_seed(Sq * 3 + Skv)
    B, heads, d = 4, 3, 64
    W = heads * d
    q = torch.randn(B * Sq, W, device="cuda").to(torch.bfloat16)
    k = torch.randn(B * Skv, W, device="cuda").to(torch.bfloat16)
    v = torch.randn(B * Skv, W, device="cuda").to(torch.bfloat16)
    dctx = torch.randn(B * Sq, W, device="cuda").to(torch.bfloat16)
    mask = torch.zeros(B, Skv, device="cuda")
    mask[1, Skv // 3:] = -10000.0
    bits = F.dropout_bits((B, heads, Sq), Skv, 0.1, 9, 0, "cuda") if drop else None
    scale = 1.0 / 0.9 if drop else 1.0
    ctx, lse2, c32 = F.attention_fwd(q, k, v, B, heads, Sq, Skv, mask, bits, scale, save_fp32=True)

    def run(flag):
        if flag:
            os.environ["MMFB_ATTN_BWD_OVERLAP"] = "1"
        else:
            os.environ.pop("MMFB_ATTN_BWD_OVERLAP", None)
        out = F.attention_bwd(dctx, q, k, v, ctx, lse2, B, heads, Sq, Skv, mask, bits, scale, ctx32=c32)
        torch.cuda.synchronize()
        return [t.clone() for t in out]
    try:
        ref, got = run(False), run(True)
    finally:
        os.environ.pop("MMFB_ATTN_BWD_OVERLAP", None)
    for r, t in zip(ref, got):
        assert torch.equal(r, t)        # the same MMAs in the same accumulation order: only the issue order differs


def test_uniter_model_base_vs_reference_golden():
    """SURVEY.md 8f item 3 (written after the round's GPU budget was spent; CPU-verified over the kernel test double)"""
    import types
    from mmf_b200.uniter import B200UNITERModelBase
    g = torch.load(os.path.join(os.path.dirname(__file__), "golden", "uniter.pt"), weights_only=False)
    c = g["cfg"]
    cfg = types.SimpleNamespace(hidden_size=c["hidden"], num_attention_heads=c["heads"], intermediate_size=c["inter"],
                                num_hidden_layers=c["layers"], vocab_size=c["vocab"], max_position_embeddings=c["max_pos"],
                                type_vocab_size=2, hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0,
                                layer_norm_eps=1e-12, initializer_range=0.02)
    m = B200UNITERModelBase(cfg, img_dim=c["img_dim"])
    m.load_state_dict({k: v for k, v in g["state_dict"].items() if k in m.state_dict()})
    m = m.cuda().eval()
    cu = lambda k: g[k].cuda()
    feat = cu("feat").requires_grad_(True)
    out = m(cu("ids"), cu("pos_ids"), feat, cu("pos"), cu("att"))
    rel = lambda a, b: ((a.double().cpu() - b.double()).norm() / b.double().norm()).item()
    assert rel(out.final_layer, g["final"]) < 1e-2 and rel(out.hidden_layers[1], g["hidden_1"]) < 1e-2
    (out.final_layer * cu("w_rand")).sum().backward()
    assert rel(feat.grad, g["dfeat"]) < 3e-2
    with torch.no_grad():      # [1, T] position ids broadcast over the batch like HF BertEmbeddings (uniter.py:732-737)
        bc = m(cu("ids"), cu("pos_ids")[:1], cu("feat"), cu("pos"), cu("att")).final_layer
        assert torch.equal(bc, m(cu("ids"), cu("pos_ids"), cu("feat"), cu("pos"), cu("att")).final_layer)


def test_lxmert_encoder_vs_reference_golden():
    """cross-modality layers with the shared cross-attention block (CPU-verified over the test double)"""
    import types
    from mmf_b200.lxmert import B200LXMERTEncoder
    g = torch.load(os.path.join(os.path.dirname(__file__), "golden", "lxmert.pt"), weights_only=False)
    c = g["cfg"]
    cfg = types.SimpleNamespace(hidden_size=c["hidden"], num_attention_heads=c["heads"], intermediate_size=c["inter"],
                                hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0, visual_feat_dim=c["feat_dim"],
                                visual_pos_dim=c["pos_dim"], l_layers=c["l"], x_layers=c["x"], r_layers=c["r"])
    enc = B200LXMERTEncoder(cfg)
    enc.load_state_dict(g["state_dict"])
    enc = enc.cuda().eval()
    lang = g["lang"].cuda().requires_grad_(True)
    feats = g["feats"].cuda().requires_grad_(True)
    ladd = ((1.0 - g["lmask"][:, None, None, :].float()) * -10000.0).cuda()
    vadd = ((1.0 - g["vmask"][:, None, None, :].float()) * -10000.0).cuda()
    lo, vo = enc(lang, ladd, (feats, g["boxes"].cuda()), vadd)
    rel = lambda a, b: ((a.double().cpu() - b.double()).norm() / b.double().norm()).item()
    assert rel(lo, g["lang_out"]) < 1e-2 and rel(vo, g["visn_out"]) < 1e-2
    ((lo * g["wl"].cuda()).sum() + (vo * g["wv"].cuda()).sum()).backward()
    assert rel(lang.grad, g["dlang"]) < 3e-2 and rel(feats.grad, g["dfeats"]) < 3e-2
    k = "x_layers.0.visual_attention.att.query.weight"          # shared block: gradient = sum of both directions
    assert rel(dict(enc.named_parameters())[k].grad, g["grads"][k]) < 5e-2


def test_gelu_bwd_kernel_and_masked_lm_head():
    """mmfb_gelu_bwd against the fp32 formula, then the head + loss against the HF golden"""
    import math
    import types
    from mmf_b200 import functional as F
    from mmf_b200.heads import B200BertPreTrainingHeads, masked_lm_loss
    torch.manual_seed(0)
    u = (torch.randn(1000, 72, device="cuda") * 2).to(torch.bfloat16)
    dh = torch.randn(1000, 72, device="cuda").to(torch.bfloat16)
    du = F.gelu_bwd(dh, u)
    uf = u.float()
    ref = dh.float() * (0.5 * (1 + torch.erf(uf / math.sqrt(2))) + uf * torch.exp(-0.5 * uf * uf) / math.sqrt(2 * math.pi))
    assert ((du.float() - ref).norm() / ref.norm()).item() < 5e-3
    tail = F.gelu_bwd(dh.reshape(-1)[:13].contiguous(), u.reshape(-1)[:13].contiguous())      # scalar tail path
    assert ((tail.float() - ref.reshape(-1)[:13]).abs().max() < 2e-2)
    g = torch.load(os.path.join(os.path.dirname(__file__), "golden", "mlm_head.pt"), weights_only=False)
    cfg = types.SimpleNamespace(hidden_size=g["cfg"]["hidden"], vocab_size=g["cfg"]["vocab"], layer_norm_eps=1e-12,
                                initializer_range=0.02)
    cls = B200BertPreTrainingHeads(cfg)
    sd = {k[len("cls."):]: v for k, v in g["state_dict"].items()}
    sd["predictions.decoder.bias"] = sd["predictions.bias"]
    cls.load_state_dict(sd)
    cls = cls.cuda().eval()
    seq = g["seq"].cuda().requires_grad_(True)
    loss, logits = masked_lm_loss(cls, seq, g["labels"].cuda())
    rel = lambda a, b: ((a.double().cpu() - b.double()).norm() / b.double().norm()).item()
    assert rel(logits, g["scores"]) < 1e-2 and abs(loss.item() - g["loss"].item()) < 1e-2 * abs(g["loss"].item())
    loss.backward()
    assert rel(seq.grad, g["dseq"]) < 3e-2


def test_visual_bert_bypass_transformer_vs_reference_golden():
    import types
    from mmf_b200.visual_bert import B200VisualBERTBase
    g = torch.load(os.path.join(os.path.dirname(__file__), "golden", "visual_bert_bypass.pt"), weights_only=False)
    c = g["cfg"]
    cfg = types.SimpleNamespace(hidden_size=c["hidden"], num_attention_heads=c["heads"], intermediate_size=c["inter"],
                                num_hidden_layers=c["layers"], vocab_size=c["vocab"], max_position_embeddings=c["max_pos"],
                                type_vocab_size=2, visual_embedding_dim=c["vdim"], hidden_dropout_prob=0.0,
                                attention_probs_dropout_prob=0.0, layer_norm_eps=1e-12, hidden_act="gelu",
                                initializer_range=0.02, bypass_transformer=True)
    m = B200VisualBERTBase(cfg)
    m.load_state_dict({k: v for k, v in g["state_dict"].items() if k in m.state_dict()})
    m = m.cuda().eval()
    feats = g["feats"].cuda().requires_grad_(True)
    seq, pooled, _ = m(g["ids"].cuda(), g["att"].cuda(), g["seg"].cuda(), feats, g["vtype"].cuda())
    rel = lambda a, b: ((a.double().cpu() - b.double()).norm() / b.double().norm()).item()
    assert rel(seq, g["seq"]) < 1e-2 and rel(pooled, g["pooled"]) < 1e-2
    (seq * g["w_rand"].cuda()).sum().backward()
    assert rel(feats.grad, g["dfeats"]) < 3e-2


@pytest.mark.parametrize("Sq,Skv,drop", [(228, 228, True), (256, 130, False), (100, 256, True), (36, 36, True), (128, 128, False)])
def test_16_warp_fused_attention_backward_matches_the_default(Sq, Skv, drop):
    """MMFB_ATTN_BWD=16: four threads per query row, overlapped issue order, tiles on four barriers - same MMAs in the same
    accumulation order and the same per-element arithmetic as the default fused kernel"""
    from mmf_b200 import functional as F
    torch.manual_seed(Sq * 5 + Skv)
    B, heads, d = 4, 3, 64
    W = heads * d
    q = torch.randn(B * Sq, W, device="cuda").to(torch.bfloat16)
    k = torch.randn(B * Skv, W, device="cuda").to(torch.bfloat16)
    v = torch.randn(B * Skv, W, device="cuda").to(torch.bfloat16)
    dctx = torch.randn(B * Sq, W, device="cuda").to(torch.bfloat16)
    mask = torch.zeros(B, Skv, device="cuda")
    mask[1, Skv // 3:] = -10000.0
    mask[2, :] = -10000.0
    bits = F.dropout_bits((B, heads, Sq), Skv, 0.1, 9, 0, "cuda") if drop else None
    scale = 1.0 / 0.9 if drop else 1.0
    ctx, lse2, c32 = F.attention_fwd(q, k, v, B, heads, Sq, Skv, mask, bits, scale, save_fp32=True)

    def run(flag):
        if flag:
            os.environ["MMFB_ATTN_BWD"] = "16"
        else:
            os.environ.pop("MMFB_ATTN_BWD", None)
        out = F.attention_bwd(dctx, q, k, v, ctx, lse2, B, heads, Sq, Skv, mask, bits, scale, ctx32=c32)
        torch.cuda.synchronize()
        return [t.clone() for t in out]
    try:
        ref, got = run(False), run(True)
    finally:
        os.environ.pop("MMFB_ATTN_BWD", None)
    for name, r, t in zip(("dq", "dk", "dv"), ref, got):
        r, t = r.float(), t.float()
        assert (r - t).abs().max() <= 1e-2 * r.abs().max(), name
        assert (r != t).float().mean() < 0.01, (name, (r != t).float().mean().item())
