"""Parity at DEPTH and per-layer gradient bars (VERDICT r1 "tighten and deepen parity").

* per layer (L = 1, config-2 / config-5 widths): output, input gradient and EVERY parameter gradient within 1e-2 relative L2
  of the fp32 oracle (the 1e-2 bf16 bar applies per layer).
* full depth (config 2: 12 layers; config 3: 12 text + 6 image + 6 co-attention layers at the real widths; config 5: 24
  layers / 1024): end to end the product must be no worse than the REFERENCE ARITHMETIC ITSELF run in bf16 (the oracle with
  bf16 weights and activations on the same device): error(product vs fp32 oracle) <= 1.5 x error(bf16 oracle vs fp32 oracle),
  with an absolute ceiling of 2.5e-2 on the outputs.
Every measured error is printed (pytest -rP keeps it; the GPU log of the round is copied to profiles/)."""
import types

import pytest
import torch

from oracle import fusion_oracle as O

pytestmark = pytest.mark.gpu


def rel(a, b):
    a, b = a.detach().double(), b.detach().double()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


def bert_cfg(hidden, heads, inter, layers):
    return types.SimpleNamespace(hidden_size=hidden, num_attention_heads=heads, intermediate_size=inter,
                                 num_hidden_layers=layers, hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0,
                                 layer_norm_eps=1e-12, hidden_act="gelu", initializer_range=0.02)


def _perturb_1d(module):
    with torch.no_grad():
        for p in module.parameters():
            if p.dim() == 1:
                p.add_(torch.randn_like(p) * 0.02)


def _oracle_bert(sd, x, add, L, heads, w, dtype):
    sdc = {k: v.detach().to(dtype).clone().requires_grad_(True) for k, v in sd.items()}
    xc = x.detach().to(dtype).clone().requires_grad_(True)
    out = O.bert_encoder(xc, add.to(dtype), sdc, "", L, heads)
    (out.float() * w).sum().backward()
    return out.detach().float(), xc.grad.float(), {k: v.grad.float() for k, v in sdc.items() if v.grad is not None}


def _key_bias(n):
    return ".key.bias" in n or ".key1.bias" in n or ".key2.bias" in n


@pytest.mark.parametrize("B,S,H,heads,I", [(4, 228, 768, 12, 3072), (4, 120, 1024, 16, 4096)])
def test_single_layer_every_gradient_within_1e2(B, S, H, heads, I):
    from mmf_b200.modules import B200BertEncoder
    torch.manual_seed(11)
    enc = B200BertEncoder(bert_cfg(H, heads, I, 1)).cuda().eval()
    _perturb_1d(enc)
    g = torch.Generator(device="cuda").manual_seed(12)
    x = torch.randn(B, S, H, generator=g, device="cuda")
    lens = torch.randint(S // 2, S + 1, (B,), generator=g, device="cuda")
    add = O.extended_attention_mask((torch.arange(S, device="cuda")[None] < lens[:, None]).long())
    w = torch.randn(B, S, H, generator=g, device="cuda")
    xg = x.clone().requires_grad_(True)
    out = enc(xg, add)[0]
    (out * w).sum().backward()
    sd = {k: v.to(torch.bfloat16).float() for k, v in enc.state_dict().items()}
    o_out, o_dx, o_g = _oracle_bert(sd, x.to(torch.bfloat16).float(), add, 1, heads, w, torch.float32)
    errs = {"out": rel(out, o_out), "dx": rel(xg.grad, o_dx)}
    named = dict(enc.named_parameters())
    for n, p in named.items():
        if _key_bias(n):        # analytically zero gradient (softmax shift invariance): only required to be tiny
            assert p.grad.float().norm() <= 0.05 * named[n.replace("key", "query")].grad.float().norm() + 1e-3, n
            continue
        errs[n] = rel(p.grad, o_g[n])
    worst = max(errs, key=errs.get)
    print("L=1 H=%d S=%d: out %.2e dx %.2e worst %s %.2e | all: %s" % (
        H, S, errs["out"], errs["dx"], worst, errs[worst], " ".join("%s=%.1e" % (k.replace("layer.0.", ""), v) for k, v in errs.items())))
    assert errs[worst] < 1e-2, (worst, errs[worst])


@pytest.mark.parametrize("name,B,S,H,heads,I,L", [("config 2 (VisualBERT 12L)", 2, 228, 768, 12, 3072, 12),
                                                  ("config 5 (24L/1024)", 2, 120, 1024, 16, 4096, 24)])
def test_full_depth_bert_stack_no_worse_than_bf16_reference(name, B, S, H, heads, I, L):
    from mmf_b200.modules import B200BertEncoder
    torch.manual_seed(21)
    enc = B200BertEncoder(bert_cfg(H, heads, I, L)).cuda().eval()
    _perturb_1d(enc)
    g = torch.Generator(device="cuda").manual_seed(22)
    x = torch.randn(B, S, H, generator=g, device="cuda")
    lens = torch.randint(S // 2, S + 1, (B,), generator=g, device="cuda")
    add = O.extended_attention_mask((torch.arange(S, device="cuda")[None] < lens[:, None]).long())
    w = torch.randn(B, S, H, generator=g, device="cuda")
    xg = x.clone().requires_grad_(True)
    out = enc(xg, add)[0]
    (out * w).sum().backward()
    sd = {k: v.to(torch.bfloat16).float() for k, v in enc.state_dict().items()}
    xr = x.to(torch.bfloat16).float()
    f_out, f_dx, f_g = _oracle_bert(sd, xr, add, L, heads, w, torch.float32)
    b_out, b_dx, b_g = _oracle_bert(sd, xr, add, L, heads, w, torch.bfloat16)       # the reference arithmetic in bf16
    probes = ["layer.0.attention.self.query.weight", "layer.0.intermediate.dense.weight",
              "layer.%d.output.dense.weight" % (L - 1), "layer.%d.attention.output.LayerNorm.weight" % (L // 2)]
    named = dict(enc.named_parameters())
    ours = {"out": rel(out, f_out), "dx": rel(xg.grad, f_dx)}
    drift = {"out": rel(b_out, f_out), "dx": rel(b_dx, f_dx)}
    for n in probes:
        ours[n], drift[n] = rel(named[n].grad, f_g[n]), rel(b_g[n], f_g[n])
    print("%s: product vs fp32 oracle / bf16-reference drift: %s" % (
        name, " ".join("%s=%.2e/%.2e" % (k.split("layer.")[-1], ours[k], drift[k]) for k in ours)))
    assert ours["out"] < 2.5e-2
    for k in ours:
        assert ours[k] <= 1.5 * drift[k] + 2e-3, (k, ours[k], drift[k])


def test_full_depth_vilbert_config3_no_worse_than_bf16_reference():
    """12 text + 6 image + 6 connection layers at the real widths (768/12h/3072, 1024/8h/1024, bi 1024/8h), T = R = 36"""
    from mmf_b200.modules import B200ViLBertEncoder
    c = dict(hidden_size=768, num_attention_heads=12, intermediate_size=3072, num_hidden_layers=12, v_hidden_size=1024,
             v_num_attention_heads=8, v_intermediate_size=1024, v_num_hidden_layers=6, bi_hidden_size=1024,
             bi_num_attention_heads=8, v_biattention_id=[0, 1, 2, 3, 4, 5], t_biattention_id=[6, 7, 8, 9, 10, 11])
    cfg = types.SimpleNamespace(hidden_dropout_prob=0.0, v_hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0,
                                v_attention_probs_dropout_prob=0.0, layer_norm_eps=1e-12, initializer_range=0.02, **c)
    torch.manual_seed(31)
    enc = B200ViLBertEncoder(cfg).cuda().eval()
    _perturb_1d(enc)
    B, T, R = 2, 36, 36
    g = torch.Generator(device="cuda").manual_seed(32)
    txt = torch.randn(B, T, 768, generator=g, device="cuda")
    img = torch.randn(B, R, 1024, generator=g, device="cuda")
    tmask = torch.ones(B, T, dtype=torch.long, device="cuda")
    tmask[1, 20:] = 0
    imask = torch.ones(B, R, dtype=torch.long, device="cuda")
    imask[0, 30:] = 0
    tadd, iadd = O.extended_attention_mask(tmask), O.extended_attention_mask(imask)
    wt = torch.randn(B, T, 768, generator=g, device="cuda")
    wv = torch.randn(B, R, 1024, generator=g, device="cuda")
    tg, ig = txt.clone().requires_grad_(True), img.clone().requires_grad_(True)
    tl, vl, _ = enc(tg, ig, tadd, tadd, iadd, None, output_all_encoded_layers=False)
    ((tl[-1] * wt).sum() + (vl[-1] * wv).sum()).backward()
    sd = {k: v.to(torch.bfloat16).float() for k, v in enc.state_dict().items()}

    def run(dtype):
        sdc = {k: v.detach().to(dtype).clone().requires_grad_(True) for k, v in sd.items()}
        tc = txt.to(torch.bfloat16).to(dtype).clone().requires_grad_(True)
        ic = img.to(torch.bfloat16).to(dtype).clone().requires_grad_(True)
        to, vo = O.vilbert_encoder(tc, ic, tadd.to(dtype), iadd.to(dtype), sdc, "", c)
        ((to.float() * wt).sum() + (vo.float() * wv).sum()).backward()
        return to.detach().float(), vo.detach().float(), tc.grad.float(), ic.grad.float(), {
            k: v.grad.float() for k, v in sdc.items() if v.grad is not None}
    f, b = run(torch.float32), run(torch.bfloat16)
    named = dict(enc.named_parameters())
    probes = ["layer.0.attention.self.query.weight", "v_layer.0.intermediate.dense.weight",
              "c_layer.0.biattention.query1.weight", "c_layer.5.t_output.dense.weight", "layer.11.output.dense.weight"]
    ours = {"t_out": rel(tl[-1], f[0]), "v_out": rel(vl[-1], f[1]), "dtxt": rel(tg.grad, f[2]), "dimg": rel(ig.grad, f[3])}
    drift = {"t_out": rel(b[0], f[0]), "v_out": rel(b[1], f[1]), "dtxt": rel(b[2], f[2]), "dimg": rel(b[3], f[3])}
    for n in probes:
        ours[n], drift[n] = rel(named[n].grad, f[4][n]), rel(b[4][n], f[4][n])
    print("config 3 (ViLBERT 12+6+6): product vs fp32 oracle / bf16-reference drift: %s" % (
        " ".join("%s=%.2e/%.2e" % (k, ours[k], drift[k]) for k in ours)))
    assert ours["t_out"] < 2.5e-2 and ours["v_out"] < 2.5e-2
    for k in ours:
        assert ours[k] <= 1.5 * drift[k] + 2e-3, (k, ours[k], drift[k])
