"""Drop-in encoder modules: same constructor inputs, sub-module / parameter names, forward signatures and
return structure as the reference's encoders, computing on the B200 engine.

  B200BertEncoder     <->  BertEncoderJit           mmf/modules/hf_layers.py:295-355
  B200ViLBertEncoder  <->  vilbert.BertEncoder      mmf/models/vilbert.py:559-796

Parameter names are the reference's (`layer.{i}.attention.self.query.weight`, `c_layer.{i}.biattention.query1.bias`,
`c_layer.{i}.biOutput.q_dense1.weight` ... SURVEY.md 5 "Checkpoint / resume"), so zoo checkpoints and
`pretrained_state_mapping` keep loading.  The nn.Linear / nn.LayerNorm children below are parameter HOLDERS:
their own forward is never called.

The same runners can be attached to an EXISTING HuggingFace / MMF module tree (mmf_b200.patch), which is how
VisualBERT / MMBT / MMFTransformer pick the engine up without any registry indirection.
"""
import os

import torch
from torch import nn

from . import engine as E
from . import lib


def _require_cuda(t, what):
    if not t.is_cuda:
        raise RuntimeError(
            "%s is on %s: the B200 fusion path has no CPU fallback (move the model and inputs to a B200)" % (what, t.device))
    if not lib.device_ok():
        raise RuntimeError("libmmfb200 needs an sm_100 (B200) device: " + lib.last_error())


def _check_erf_gelu(layer):
    """The FFN epilogue IS the exact erf GELU of the reference configs (ACT2FN["gelu"]); an attached HuggingFace layer
    with any other activation (gelu_new, relu, swish ...) must not be rerouted silently."""
    fn = getattr(getattr(layer, "intermediate", None), "intermediate_act_fn", None)
    if fn is None:
        return
    name = getattr(fn, "__name__", None) or type(fn).__name__
    if name.lower() in ("gelu", "geluactivation"):
        return
    probe = torch.tensor([-1.5, -0.3, 0.7, 2.0])
    try:
        ok = torch.allclose(fn(probe), torch.nn.functional.gelu(probe), atol=1e-6)
    except Exception:
        ok = False
    if not ok:
        raise NotImplementedError("the B200 fusion block implements the erf GELU only; this encoder uses %s" % name)


_SEED_COUNTER = [0]


def _fresh_dropout_state(prefetchable=False):
    """One Philox stream per forward, derived from torch's seed so `torch.manual_seed` makes runs repeatable.
    (A side-stream generator that drew the next layer's bits ahead was measured in round 2 and dropped: +0.2 %, i.e. noise -
    the persistent GEMMs leave no SM resources for a concurrent kernel; profiles/r2_ab_sweep_first.txt.)"""
    _SEED_COUNTER[0] += 1
    seed = torch.initial_seed() * 1000003 + _SEED_COUNTER[0]
    return E.DropoutState(seed)


# ------------------------------------------------------------------------------------------------------
# single-stream BERT encoder
# ------------------------------------------------------------------------------------------------------
class EncoderRunner:
    """Engine state of one BERT-style layer stack (built lazily on the first CUDA forward)."""

    W = E.BertLayerW                       # parameter handles of one layer
    _fwd = staticmethod(E.bert_layer_fwd)  # (x, add_mask, w, B, S, p_attn, p_hidden, ds) -> (out, saved)
    _bwd = staticmethod(E.bert_layer_bwd)  # (dout, saved, add_mask, w, B, S) -> dx

    def __init__(self, layers):
        self.layers = list(layers)
        self.pack = None
        self.weights = None
        self.grad_ready_hook = None   # callable(layer_index) - set by mmf_b200.ddp for bucket all-reduce overlap

    def _probs(self, m):
        def p_of(mod):
            d = getattr(mod, "dropout", None)
            return float(getattr(d, "p", 0.0)) if d is not None else 0.0
        return p_of(m.attention.self), p_of(m.attention.output)

    def ensure(self, device):
        if self.pack is not None and self.pack.intact() and self.pack.device == device:
            return
        params = []
        for m in self.layers:
            _check_erf_gelu(m)
            params += self.W.params(m)
        for p in params:
            _require_cuda(p, "encoder parameter")
        self.pack = E.ParamPack(params, device)
        self.weights = [self.W(self.pack, m) for m in self.layers]
        self.layer_param_ranges = []
        per = len(self.W.params(self.layers[0])) if self.layers else 0
        for i in range(len(self.layers)):
            o0 = self.pack.offsets[i * per]
            o1 = self.pack.offsets[(i + 1) * per] if (i + 1) * per < len(self.pack.offsets) else self.pack.total
            self.layer_param_ranges.append((o0, o1))

    def forward(self, x, add_mask, training, need_grad, first_layer=0, last_layer=None):
        B, S, H = x.shape
        h = x.reshape(B * S, H)
        saved = []
        ds = _fresh_dropout_state(prefetchable=True) if training else None
        last_layer = len(self.layers) if last_layer is None else last_layer
        hiddens = []
        for i in range(first_layer, last_layer):
            m = self.layers[i]
            pa, ph = self._probs(m) if training else (0.0, 0.0)
            hiddens.append(h)
            h, s = self._fwd(h, add_mask, self.weights[i], B, S, pa, ph, ds)
            saved.append(s if need_grad else None)
        return h, saved, hiddens

    def backward(self, dout, saved, add_mask, B, S, first_layer=0):
        d = dout
        for j in range(len(saved) - 1, -1, -1):
            i = first_layer + j
            d = self._bwd(d, saved[j], add_mask, self.weights[i], B, S)
            saved[j] = None
            if self.grad_ready_hook is not None:
                E.join_side(d.device)
                self.grad_ready_hook(i)
        return d


class _BertEncoderFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, runner, add_mask, training, want_hiddens, x, *params):
        B, S, H = x.shape
        need_grad = any(ctx.needs_input_grad)   # (grad mode itself is always off inside Function.forward)
        runner.pack.refresh()
        xb = x.detach().to(torch.bfloat16).contiguous()
        out, saved, hiddens = runner.forward(xb, add_mask, training, need_grad)
        ctx.runner, ctx.saved, ctx.add_mask, ctx.dims, ctx.in_dtype = runner, saved, add_mask, (B, S, H), x.dtype
        outs = [out.view(B, S, H).to(x.dtype)]
        if want_hiddens:
            outs += [h.view(B, S, H).to(x.dtype) for h in hiddens]
            ctx.mark_non_differentiable(*outs[1:])
        return tuple(outs)

    @staticmethod
    def backward(ctx, dout, *unused):
        runner = ctx.runner
        B, S, H = ctx.dims
        if ctx.saved is None or any(s is None for s in ctx.saved):
            raise RuntimeError("B200 encoder: backward called twice or activations were not saved")
        aliased = runner.pack.prepare_grads()
        d = dout.to(torch.bfloat16).contiguous().view(B * S, H)
        dx = runner.backward(d, ctx.saved, ctx.add_mask, B, S)
        ctx.saved = None
        return (None, None, None, None, dx.view(B, S, H).to(ctx.in_dtype)) + tuple(runner.pack.autograd_grads(aliased))


def run_bert_encoder(runner, hidden_states, attention_mask, training, output_hidden_states=False):
    """hidden_states [B,S,H] (fp32 or bf16), attention_mask additive [B,1,1,S] / [B,S] / None."""
    _require_cuda(hidden_states, "hidden_states")
    if hidden_states.dim() != 3:
        raise ValueError("hidden_states must be [batch, seq, hidden], got %s" % (tuple(hidden_states.shape),))
    B, S, H = hidden_states.shape
    runner.ensure(hidden_states.device)
    if H != runner.weights[0].hidden:
        raise ValueError("hidden size %d does not match the encoder's %d" % (H, runner.weights[0].hidden))
    mask = E.additive_mask_2d(attention_mask, B, S)
    outs = _BertEncoderFn.apply(runner, mask, training, output_hidden_states, hidden_states, *runner.pack.params)
    return outs


class _Holder(nn.Module):
    def forward(self, *a, **k):  # pragma: no cover
        raise RuntimeError("parameter holder: computed by the B200 engine, not callable")


class B200BertSelfAttention(_Holder):
    def __init__(self, hidden, heads, p_attn, in_hidden=None):
        super().__init__()
        if hidden % heads != 0:
            raise ValueError("The hidden size (%d) is not a multiple of the number of attention heads (%d)" % (hidden, heads))
        self.num_attention_heads = heads
        self.attention_head_size = hidden // heads
        self.all_head_size = hidden
        self.query = nn.Linear(in_hidden or hidden, hidden)
        self.key = nn.Linear(in_hidden or hidden, hidden)
        self.value = nn.Linear(in_hidden or hidden, hidden)
        self.dropout = nn.Dropout(p_attn)


class B200BertSelfOutput(_Holder):
    def __init__(self, in_features, hidden, p_hidden, eps):
        super().__init__()
        self.dense = nn.Linear(in_features, hidden)
        self.LayerNorm = nn.LayerNorm(hidden, eps=eps)
        self.dropout = nn.Dropout(p_hidden)


class B200BertAttention(_Holder):
    def __init__(self, hidden, heads, p_attn, p_hidden, eps):
        super().__init__()
        self.self = B200BertSelfAttention(hidden, heads, p_attn)
        self.output = B200BertSelfOutput(hidden, hidden, p_hidden, eps)


class B200BertIntermediate(_Holder):
    def __init__(self, hidden, inter):
        super().__init__()
        self.dense = nn.Linear(hidden, inter)


class B200BertLayer(_Holder):
    def __init__(self, hidden, heads, inter, p_attn, p_hidden, eps):
        super().__init__()
        self.attention = B200BertAttention(hidden, heads, p_attn, p_hidden, eps)
        self.intermediate = B200BertIntermediate(hidden, inter)
        self.output = B200BertSelfOutput(inter, hidden, p_hidden, eps)


def _init_bert_weights(module, std):
    """normal(0, initializer_range) Linear/Embedding weights, zero biases, LayerNorm (1, 0):
    mmf/models/transformers/base.py:213-223 / HF _init_weights."""
    for m in module.modules():
        if isinstance(m, (nn.Linear, nn.Embedding)):
            m.weight.data.normal_(mean=0.0, std=std)
            if isinstance(m, nn.Linear) and m.bias is not None:
                m.bias.data.zero_()
        elif isinstance(m, nn.LayerNorm):
            m.weight.data.fill_(1.0)
            m.bias.data.zero_()


class B200BertEncoder(nn.Module):
    """BertEncoderJit with the B200 engine underneath.  `config`: a HF BertConfig-like object."""

    def __init__(self, config):
        super().__init__()
        self.config = config
        eps = float(getattr(config, "layer_norm_eps", 1e-12))
        act = getattr(config, "hidden_act", "gelu")
        if act not in ("gelu",) and not callable(act):
            raise ValueError("B200 fusion block implements the erf GELU of the reference configs, got %r" % (act,))
        self.layer = nn.ModuleList([
            B200BertLayer(config.hidden_size, config.num_attention_heads, config.intermediate_size,
                          float(config.attention_probs_dropout_prob), float(config.hidden_dropout_prob), eps)
            for _ in range(config.num_hidden_layers)])
        _init_bert_weights(self, float(getattr(config, "initializer_range", 0.02)))
        self._runner = EncoderRunner(self.layer)

    def forward(self, hidden_states, attention_mask=None, encoder_hidden_states=None, encoder_attention_mask=None,
                output_attentions=False, output_hidden_states=False, return_dict=False, head_mask=None):
        """Same contract as BertEncoderJit.forward (hf_layers.py:316-355): returns (last_hidden,
        [all_hidden_states], ...).  Attention probabilities never exist in HBM on this path."""
        if output_attentions:
            raise NotImplementedError("output_attentions: the fused kernel never materialises attention probabilities")
        if head_mask is not None and not (isinstance(head_mask, (list, tuple)) and all(h is None for h in head_mask)):
            raise NotImplementedError("head_mask is not supported on the B200 path")
        # `encoder_hidden_states` is ignored exactly as BertLayerJit ignores it (hf_layers.py:281-283; MMFT passes
        # `[None]*L` there, huggingface.py:229-235)
        outs = run_bert_encoder(self._runner, hidden_states, attention_mask, self.training, output_hidden_states)
        if output_hidden_states:
            return (outs[0], tuple(outs[1:]) + (outs[0],))
        return (outs[0],)


# ------------------------------------------------------------------------------------------------------
# ViLBERT two-stream encoder
# ------------------------------------------------------------------------------------------------------
class B200BiAttention(_Holder):
    def __init__(self, cfg):
        super().__init__()
        bi, hv, ht = cfg.bi_hidden_size, cfg.v_hidden_size, cfg.hidden_size
        if bi % cfg.bi_num_attention_heads != 0:
            raise ValueError("The hidden size (%d) is not a multiple of the number of attention heads (%d)"
                             % (bi, cfg.bi_num_attention_heads))
        self.num_attention_heads = cfg.bi_num_attention_heads
        self.attention_head_size = bi // cfg.bi_num_attention_heads
        self.all_head_size = bi
        self.query1, self.key1, self.value1 = nn.Linear(hv, bi), nn.Linear(hv, bi), nn.Linear(hv, bi)
        self.dropout1 = nn.Dropout(cfg.v_attention_probs_dropout_prob)
        self.query2, self.key2, self.value2 = nn.Linear(ht, bi), nn.Linear(ht, bi), nn.Linear(ht, bi)
        self.dropout2 = nn.Dropout(cfg.attention_probs_dropout_prob)


class B200BiOutput(_Holder):
    def __init__(self, cfg):
        super().__init__()
        bi, hv, ht = cfg.bi_hidden_size, cfg.v_hidden_size, cfg.hidden_size
        self.dense1 = nn.Linear(bi, hv)
        self.LayerNorm1 = nn.LayerNorm(hv, eps=1e-12)
        self.dropout1 = nn.Dropout(cfg.v_hidden_dropout_prob)
        self.q_dense1 = nn.Linear(bi, hv)      # present in the reference, never used (vilbert.py:486-494): no grads
        self.q_dropout1 = nn.Dropout(cfg.v_hidden_dropout_prob)
        self.dense2 = nn.Linear(bi, ht)
        self.LayerNorm2 = nn.LayerNorm(ht, eps=1e-12)
        self.dropout2 = nn.Dropout(cfg.hidden_dropout_prob)
        self.q_dense2 = nn.Linear(bi, ht)
        self.q_dropout2 = nn.Dropout(cfg.hidden_dropout_prob)


class B200ConnectionLayer(_Holder):
    def __init__(self, cfg):
        super().__init__()
        self.biattention = B200BiAttention(cfg)
        self.biOutput = B200BiOutput(cfg)
        self.v_intermediate = B200BertIntermediate(cfg.v_hidden_size, cfg.v_intermediate_size)
        self.v_output = B200BertSelfOutput(cfg.v_intermediate_size, cfg.v_hidden_size, cfg.v_hidden_dropout_prob, 1e-12)
        self.t_intermediate = B200BertIntermediate(cfg.hidden_size, cfg.intermediate_size)
        self.t_output = B200BertSelfOutput(cfg.intermediate_size, cfg.hidden_size, cfg.hidden_dropout_prob, 1e-12)


def vilbert_schedule(v_ids, t_ids, n_t, n_v):
    """Layer interleaving of vilbert.BertEncoder.forward (vilbert.py:619-785) as ('t'|'v'|'c', index) steps."""
    steps, v_start, t_start = [], 0, 0
    for count, (v_end, t_end) in enumerate(zip(v_ids, t_ids)):
        steps += [("t", i) for i in range(t_start, t_end)]
        steps += [("v", i) for i in range(v_start, v_end)]
        steps.append(("c", count))
        v_start, t_start = v_end, t_end
    steps += [("v", i) for i in range(v_start, n_v)]
    steps += [("t", i) for i in range(t_start, n_t)]
    return steps


class VilbertRunner:
    def __init__(self, layer, v_layer, c_layer, v_ids, t_ids):
        self.layer, self.v_layer, self.c_layer = list(layer), list(v_layer), list(c_layer)
        self.steps = vilbert_schedule(list(v_ids), list(t_ids), len(self.layer), len(self.v_layer))
        self.pack = None
        self.grad_ready_hook = None

    def ensure(self, device):
        if self.pack is not None and self.pack.intact() and self.pack.device == device:
            return
        params = []
        # parameter order = execution order, so that gradient buckets complete in reverse order
        for kind, i in self.steps:
            m = {"t": self.layer, "v": self.v_layer, "c": self.c_layer}[kind][i]
            params += (E.ConnectionW if kind == "c" else E.BertLayerW).params(m)
        for p in params:
            _require_cuda(p, "encoder parameter")
        self.pack = E.ParamPack(params, device)
        self.w = {}
        for kind, i in self.steps:
            m = {"t": self.layer, "v": self.v_layer, "c": self.c_layer}[kind][i]
            self.w[(kind, i)] = (E.ConnectionW if kind == "c" else E.BertLayerW)(self.pack, m)

    @staticmethod
    def _p(mod):
        d = getattr(mod, "dropout", None)
        return float(getattr(d, "p", 0.0)) if d is not None else 0.0

    def forward(self, txt, img, tmask, imask, B, T, R, training, need_grad, block_states=None):
        ds = _fresh_dropout_state() if training else None
        saved = []
        for kind, i in self.steps:
            w = self.w[(kind, i)]
            if kind == "t":
                m = self.layer[i]
                pa, ph = (self._p(m.attention.self), self._p(m.attention.output)) if training else (0.0, 0.0)
                txt, s = E.bert_layer_fwd(txt, tmask, w, B, T, pa, ph, ds)
            elif kind == "v":
                m = self.v_layer[i]
                pa, ph = (self._p(m.attention.self), self._p(m.attention.output)) if training else (0.0, 0.0)
                img, s = E.bert_layer_fwd(img, imask, w, B, R, pa, ph, ds)
            else:
                m = self.c_layer[i]
                if training:
                    pva, pta = float(m.biattention.dropout1.p), float(m.biattention.dropout2.p)
                    pvh, pth = float(m.biOutput.dropout1.p), float(m.biOutput.dropout2.p)
                else:
                    pva = pta = pvh = pth = 0.0
                img, txt, s = E.connection_fwd(img, txt, imask, tmask, w, B, R, T, pva, pta, pvh, pth, ds)
                if block_states is not None:
                    block_states.append((txt, img))      # the states the reference appends per co-attention block
            saved.append(s if need_grad else None)
        return txt, img, saved

    def backward(self, dtxt, dimg, saved, tmask, imask, B, T, R):
        for j in range(len(self.steps) - 1, -1, -1):
            kind, i = self.steps[j]
            w = self.w[(kind, i)]
            if kind == "t":
                dtxt = E.bert_layer_bwd(dtxt, saved[j], tmask, w, B, T)
            elif kind == "v":
                dimg = E.bert_layer_bwd(dimg, saved[j], imask, w, B, R)
            else:
                dimg, dtxt = E.connection_bwd(dimg, dtxt, saved[j], imask, tmask, w, B, R, T)
            saved[j] = None
            if self.grad_ready_hook is not None:
                E.join_side()
                self.grad_ready_hook(j)
        return dtxt, dimg


class _VilbertEncoderFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, runner, tmask, imask, training, want_blocks, txt, img, *params):
        B, T, Ht = txt.shape
        _, R, Hv = img.shape
        need_grad = any(ctx.needs_input_grad)
        runner.pack.refresh()
        t = txt.detach().to(torch.bfloat16).contiguous().view(B * T, Ht)
        v = img.detach().to(torch.bfloat16).contiguous().view(B * R, Hv)
        blocks = [] if want_blocks else None
        t, v, saved = runner.forward(t, v, tmask, imask, B, T, R, training, need_grad, blocks)
        ctx.runner, ctx.saved, ctx.masks, ctx.dims, ctx.dtypes = runner, saved, (tmask, imask), (B, T, R, Ht, Hv), (txt.dtype, img.dtype)
        outs = [t.view(B, T, Ht).to(txt.dtype), v.view(B, R, Hv).to(img.dtype)]
        if want_blocks:
            for bt, bv in blocks:
                outs += [bt.view(B, T, Ht).to(txt.dtype), bv.view(B, R, Hv).to(img.dtype)]
            ctx.mark_non_differentiable(*outs[2:])
        return tuple(outs)

    @staticmethod
    def backward(ctx, dtxt, dimg, *unused):
        runner = ctx.runner
        B, T, R, Ht, Hv = ctx.dims
        if ctx.saved is None:
            raise RuntimeError("B200 ViLBERT encoder: backward called twice")
        aliased = runner.pack.prepare_grads()
        dt = (dtxt if dtxt is not None else torch.zeros(B, T, Ht, device=runner.pack.device)).to(torch.bfloat16).contiguous().view(B * T, Ht)
        dv = (dimg if dimg is not None else torch.zeros(B, R, Hv, device=runner.pack.device)).to(torch.bfloat16).contiguous().view(B * R, Hv)
        dt, dv = runner.backward(dt, dv, ctx.saved, ctx.masks[0], ctx.masks[1], B, T, R)
        ctx.saved = None
        return (None, None, None, None, None, dt.view(B, T, Ht).to(ctx.dtypes[0]), dv.view(B, R, Hv).to(ctx.dtypes[1])) + tuple(
            runner.pack.autograd_grads(aliased))


class B200ViLBertEncoder(nn.Module):
    """vilbert.BertEncoder (vilbert.py:559-796) on the B200 engine.  `config` carries the reference's ViLBERT keys
    (mmf/configs/models/vilbert/defaults.yaml).  dynamic_attention, in_batch_pairs, fast_mode and fixed_*_layer > 0
    are configuration corners the reference defaults leave off; they raise here instead of silently differing."""

    def __init__(self, config):
        super().__init__()
        for flag in ("dynamic_attention", "in_batch_pairs", "fast_mode"):
            if getattr(config, flag, False):
                raise NotImplementedError("ViLBERT option %s is not implemented on the B200 path" % flag)
        if getattr(config, "fixed_t_layer", 0) or getattr(config, "fixed_v_layer", 0):
            raise NotImplementedError("fixed_t_layer / fixed_v_layer > 0 are not implemented on the B200 path")
        if not getattr(config, "with_coattention", True):
            raise NotImplementedError("with_coattention=False is not implemented on the B200 path")
        self.config = config
        self.v_biattention_id = list(config.v_biattention_id)
        self.t_biattention_id = list(config.t_biattention_id)
        self.layer = nn.ModuleList([
            B200BertLayer(config.hidden_size, config.num_attention_heads, config.intermediate_size,
                          float(config.attention_probs_dropout_prob), float(config.hidden_dropout_prob), 1e-12)
            for _ in range(config.num_hidden_layers)])
        self.v_layer = nn.ModuleList([
            B200BertLayer(config.v_hidden_size, config.v_num_attention_heads, config.v_intermediate_size,
                          float(config.v_attention_probs_dropout_prob), float(config.v_hidden_dropout_prob), 1e-12)
            for _ in range(config.v_num_hidden_layers)])
        self.c_layer = nn.ModuleList([B200ConnectionLayer(config) for _ in range(len(self.v_biattention_id))])
        _init_bert_weights(self, float(getattr(config, "initializer_range", 0.02)))
        self._runner = VilbertRunner(self.layer, self.v_layer, self.c_layer, self.v_biattention_id, self.t_biattention_id)

    def forward(self, txt_embedding, image_embedding, txt_attention_mask, txt_attention_mask2, image_attention_mask,
                co_attention_mask=None, output_all_encoded_layers=True, output_all_attention_masks=False):
        """Same signature / return structure as vilbert.BertEncoder.forward (vilbert.py:590-796):
        ([text layers], [image layers], ([], [], [])).  co_attention_mask is accepted and, like in the reference
        (vilbert.py:424-425,448-449), not applied.  output_all_encoded_layers=True reproduces the reference's lists: ONE
        entry per co-attention block - the states right after that block - and NOT the final states (vilbert.py:761-763,
        787-790).  Those per-block tensors are outputs for inspection: they carry no gradient here (the reference would
        back-propagate through them); the differentiable result is the default output_all_encoded_layers=False path."""
        if output_all_attention_masks:
            raise NotImplementedError("attention probabilities are never materialised on the B200 path")
        _require_cuda(txt_embedding, "txt_embedding")
        B, T, _ = txt_embedding.shape
        R = image_embedding.shape[1]
        self._runner.ensure(txt_embedding.device)
        tmask = E.additive_mask_2d(txt_attention_mask, B, T)
        imask = E.additive_mask_2d(image_attention_mask, B, R)
        outs = _VilbertEncoderFn.apply(self._runner, tmask, imask, self.training, bool(output_all_encoded_layers),
                                       txt_embedding, image_embedding, *self._runner.pack.params)
        if output_all_encoded_layers:
            return list(outs[2::2]), list(outs[3::2]), ([], [], [])
        return [outs[0]], [outs[1]], ([], [], [])
