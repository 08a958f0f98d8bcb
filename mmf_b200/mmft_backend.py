"""MMFTransformer backend plugin on the B200 engine (BASELINE.json config 4).

  B200HuggingfaceEmbeddings  <->  HuggingfaceEmbeddings   mmf/models/transformers/backends/huggingface.py:19-159
  B200TransformerBackend     <->  BaseTransformerBackend / HuggingfaceBackend
                                  mmf/models/transformers/base.py:293-377, backends/huggingface.py:162-235

Registered as transformer backend "b200" in the registry shim (mmf_b200.registry); under a real MMF install the same
class body registers with `mmf.common.registry`.  Parameter names follow the reference
(`embeddings.token_embeddings.{i}`, `embeddings.pos_embeddings.{i}`, `embeddings.layer_norms.{i}`,
`embeddings.token_type_embeddings`, `transformer.encoder.layer.*`).
"""
from copy import deepcopy

import torch
from torch import nn

from . import ops
from .mmbt import _BertModelHolder
from .modules import _init_bert_weights, _require_cuda
from .registry import registry


def _get(obj, key, default=None):
    if isinstance(obj, dict):
        return obj.get(key, default)
    return getattr(obj, key, default)


class B200HuggingfaceEmbeddings(nn.Module):
    def __init__(self, model_config, transformer_config, transformer):
        super().__init__()
        self.model_config, self.transformer_config = model_config, transformer_config
        H = transformer_config.hidden_size
        self.token_embeddings = nn.ModuleList()
        self.pos_embeddings = nn.ModuleList()
        self.layer_norms = nn.ModuleList()
        self.dropouts = nn.ModuleList()
        self.modality_keys = []
        mods = _get(model_config, "modalities")
        for m in mods:
            self.modality_keys.append(_get(m, "key"))
            eps = _get(m, "layer_norm_eps", transformer_config.layer_norm_eps)
            pdim = _get(m, "position_dim", transformer_config.max_position_embeddings)
            pdrop = _get(m, "hidden_dropout_prob", transformer_config.hidden_dropout_prob)
            if _get(m, "type") == "text" and _get(m, "consume_raw", True):
                self.token_embeddings.append(nn.Embedding(transformer_config.vocab_size, H,
                                                          padding_idx=getattr(transformer_config, "pad_token_id", 0)))
            else:
                self.token_embeddings.append(nn.Sequential(nn.Linear(_get(m, "embedding_dim"), H),
                                                           nn.LayerNorm(H, eps=eps)))
            self.pos_embeddings.append(nn.Embedding(pdim, H))
            self.layer_norms.append(nn.LayerNorm(H, eps=eps))
            self.dropouts.append(nn.Dropout(pdrop))
        self.token_type_embeddings = nn.Embedding(len(mods), H)
        _init_bert_weights(self, float(getattr(transformer_config, "initializer_range", 0.02)))
        self.init_weights(transformer)

    def init_weights(self, transformer):
        """huggingface.py:103-129: text modality shares the transformer's word table and LayerNorm; position tables
        start from the transformer's; extra type rows = mean + noise."""
        mods = _get(self.model_config, "modalities")
        for idx, m in enumerate(mods):
            if _get(m, "type") == "text":
                self.token_embeddings[idx] = transformer.embeddings.word_embeddings
                self.layer_norms[idx] = transformer.embeddings.LayerNorm
            # the reference REPLACES the table with a copy of the transformer's (so its shape becomes
            # [max_position_embeddings, H] whatever position_dim said) - huggingface.py:109-112
            self.pos_embeddings[idx].weight = nn.Parameter(
                deepcopy(transformer.embeddings.position_embeddings.weight.data), requires_grad=True)
        tv = transformer.embeddings.token_type_embeddings.weight.shape[0]
        n = min(tv, len(mods))
        self.token_type_embeddings.weight.data[:n].copy_(transformer.embeddings.token_type_embeddings.weight.data[:n])
        for idx in range(tv, len(mods)):
            self.token_type_embeddings.weight.data[idx].copy_(
                transformer.embeddings.token_type_embeddings.weight.data.mean(dim=0))
            self.token_type_embeddings.weight.data[idx] += torch.normal(
                float(_get(self.model_config, "token_noise_mean", 0.0)), float(_get(self.model_config, "token_noise_std", 0.01)),
                size=self.token_type_embeddings.weight.data[idx].size())

    def forward(self, tokens_ids, position_ids, segment_ids):
        outs = []
        for idx, key in enumerate(self.modality_keys):
            tok = tokens_ids[key]
            _require_cuda(tok, "tokens_ids[%s]" % key)
            B, N = tok.shape[0], tok.shape[1]
            H = self.transformer_config.hidden_size
            ln = self.layer_norms[idx]
            srcs, tabs = [], []
            te = self.token_embeddings[idx]
            if isinstance(te, nn.Embedding):
                tabs.append((te.weight, ops.i32(tok), te.padding_idx))
            else:
                lin, ln_in = te[0], te[1]
                x = ops.linear(tok.reshape(B * N, -1), lin.weight, lin.bias)
                x = ops.layer_norm(x, ln_in.weight, ln_in.bias, ln_in.eps)
                srcs.append((x, torch.arange(B * N, device=tok.device, dtype=torch.int32)))
            if key in position_ids:
                tabs.append((self.pos_embeddings[idx].weight, ops.i32(position_ids[key])))
            if key in segment_ids:
                tabs.append((self.token_type_embeddings.weight, ops.i32(segment_ids[key])))
            y = ops.compose_ln(B * N, H, srcs, tabs, ln.weight, ln.bias, ln.eps, float(self.dropouts[idx].p),
                               self.training)
            outs.append(y.view(B, N, H))
        return torch.cat(outs, dim=1)


@registry.register_transformer_backend("b200")
class B200TransformerBackend(nn.Module):
    """config: object/dict with `modalities` (list of {type, key, embedding_dim, position_dim, segment_id, ...}) and
    `transformer_config` (BertConfig-like).  The reference resolves the latter with AutoConfig.from_pretrained
    (network); here it is passed explicitly (random init or load_state_dict)."""

    def __init__(self, config, *args, **kwargs):
        super().__init__()
        self.config = config
        self.build_transformer_config()
        self.build_transformer_base()
        self.build_embeddings()

    def build_transformer_config(self):
        self.transformer_config = _get(self.config, "transformer_config")
        if self.transformer_config is None:
            raise ValueError("B200TransformerBackend needs config.transformer_config (a BertConfig-like object)")

    def build_transformer_base(self):
        self.transformer = _BertModelHolder(self.transformer_config)
        _init_bert_weights(self.transformer, float(getattr(self.transformer_config, "initializer_range", 0.02)))

    def build_embeddings(self):
        self.embeddings = B200HuggingfaceEmbeddings(self.config, self.transformer_config, self.transformer)

    def get_config(self):
        return self.transformer_config

    def generate_embeddings(self, tokens_ids, position_ids, segment_ids, attention_mask):
        return self.embeddings(tokens_ids=tokens_ids, position_ids=position_ids, segment_ids=segment_ids)

    def generate_attention_mask(self, masks):
        """huggingface.py:216-222"""
        attention_mask = torch.cat(masks, dim=-1)
        return (1.0 - attention_mask.unsqueeze(1).unsqueeze(2)) * -10000.0

    def generate_encoded_layers(self, embedding, attention_mask):
        """huggingface.py:224-235: returns (encoded_layers[-1], encoded_layers[0]) - the same tensor when hidden
        states are not requested."""
        enc = self.transformer.encoder(embedding, attention_mask, [None] * len(self.transformer.encoder.layer))
        return enc[-1], enc[0]

    def forward(self, tokens_ids, position_ids, segment_ids, masks):
        """BaseTransformerBackend.forward, base.py:358-377 -> (sequence_output, encoded_layers)"""
        attention_mask = self.generate_attention_mask(masks)
        embedding = self.generate_embeddings(tokens_ids, position_ids, segment_ids, attention_mask)
        return self.generate_encoded_layers(embedding, attention_mask)
