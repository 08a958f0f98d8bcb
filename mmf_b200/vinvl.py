"""VinVL (Oscar-style) trunk on the B200 engine.

  B200VinVLBase  <->  VinVLBase   mmf/models/vinvl.py:43-122: BERT text embeddings ; region features [B, R, 2054] through
                      `img_embedding` = Linear(img_feature_dim -> H) [-> LayerNorm] -> Dropout ; concatenation ; BERT encoder
                      with all hidden states returned as `TransformerOutput(last_hidden_state, hidden_layers)`.

Parameter names are the reference's (`embeddings.*`, `encoder.layer.*`, `img_embedding.0.*` projection, `img_embedding.1.*`
LayerNorm).  img_feature_dim = 2054 is not a multiple of 8 (the GEMM's 16-byte row pitch): ops.linear_any zero-pads the
feature columns and the weight.  A composition of existing kernels - no new device code.
"""
from collections import namedtuple

import torch
from torch import nn

from . import ops
from .modules import B200BertEncoder, _init_bert_weights, _require_cuda
from .vilbert import B200BertTextEmbeddings


class B200VinVLBase(nn.Module):
    def __init__(self, config):
        """config: BertConfig-like + img_feature_dim, use_img_layernorm, img_layer_norm_eps (vinvl.py:125-147)"""
        super().__init__()
        self.config = config
        self.embeddings = B200BertTextEmbeddings(config)
        self.encoder = B200BertEncoder(config)
        self.img_dim = config.img_feature_dim
        self.use_img_layernorm = bool(getattr(config, "use_img_layernorm", False))
        mods = [nn.Linear(self.img_dim, config.hidden_size, bias=True)]
        if self.use_img_layernorm:
            mods.append(nn.LayerNorm(config.hidden_size, eps=float(getattr(config, "img_layer_norm_eps", 1e-12))))
        mods.append(nn.Dropout(float(config.hidden_dropout_prob)))
        self.img_embedding = nn.Sequential(*mods)           # holder: computed below on the kernels
        _init_bert_weights(self, float(getattr(config, "initializer_range", 0.02)))

    def _img_embed(self, img_feats):
        B, R, Fd = img_feats.shape
        lin = self.img_embedding[0]
        y = ops.linear_any(img_feats.reshape(B * R, Fd), lin.weight, lin.bias)
        if self.use_img_layernorm:
            ln = self.img_embedding[1]
            y = ops.layer_norm(y, ln.weight, ln.bias, ln.eps)
        return self.img_embedding[-1](y.view(B, R, -1).to(lin.weight.dtype))     # nn.Dropout on the [B, R, H] result

    def forward(self, input_ids, img_feats, token_type_ids=None, attention_mask=None, position_ids=None):
        _require_cuda(input_ids, "input_ids")
        if attention_mask is None:
            attention_mask = torch.ones((input_ids.size(0), input_ids.size(1) + img_feats.size(1)), device=input_ids.device)
        if token_type_ids is None:
            token_type_ids = torch.zeros_like(input_ids)
        if attention_mask.dim() == 3:
            raise NotImplementedError("per-query attention masks [B, S, S] are not on the B200 path (key-padding masks only)")
        if attention_mask.dim() != 2:
            raise ValueError("Wrong shape for input_ids (shape %s) or attention_mask (shape %s)"
                             % (tuple(input_ids.shape), tuple(attention_mask.shape)))
        dt = self.img_embedding[0].weight.dtype
        ext = (1.0 - attention_mask[:, None, None, :].to(dtype=dt)) * -10000.0
        text = self.embeddings(input_ids, token_type_ids=token_type_ids, position_ids=position_ids)
        emb = torch.cat((text.to(dt), self._img_embed(img_feats)), 1)
        out = self.encoder(emb, ext, output_hidden_states=True)
        layers = namedtuple("TransformerOutput", ["last_hidden_state", "hidden_layers"])
        return layers(out[0].to(dt), out[1])
