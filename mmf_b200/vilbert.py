"""ViLBERT front-end pieces on the B200 engine (BASELINE.json config 3).

  B200ImageFeatureEmbeddings  <->  BertImageFeatureEmbeddings   mmf/models/vilbert.py:891-913
  B200ViLBERTBase             <->  ViLBERTBase.forward (masks + embeddings + two-stream encoder)  vilbert.py:916-1051
"""
import torch
from torch import nn

from . import ops
from .mmbt import _BertEmbeddingsHolder
from .modules import B200ViLBertEncoder, _init_bert_weights, _require_cuda


class B200ImageFeatureEmbeddings(nn.Module):
    """LN(Linear(v_feature_size -> v_hidden)(feat) + Linear(5 -> v_hidden)(loc)) -> dropout."""

    def __init__(self, config):
        super().__init__()
        self.image_embeddings = nn.Linear(config.v_feature_size, config.v_hidden_size)
        self.image_location_embeddings = nn.Linear(5, config.v_hidden_size)
        self.LayerNorm = nn.LayerNorm(config.v_hidden_size, eps=1e-12)
        self.dropout = nn.Dropout(float(config.hidden_dropout_prob))

    def forward(self, image_feature, image_location):
        _require_cuda(image_feature, "image_feature")
        B, R, Fd = image_feature.shape
        Hv = self.LayerNorm.weight.shape[0]
        img = ops.linear(image_feature.reshape(B * R, Fd), self.image_embeddings.weight, self.image_embeddings.bias)
        # K = 5 is below the 16-byte TMA row pitch: zero-pad the location columns (and the weight) to 8
        loc8 = nn.functional.pad(image_location.reshape(B * R, 5).to(self.image_location_embeddings.weight.dtype), (0, 3))
        w8 = nn.functional.pad(self.image_location_embeddings.weight, (0, 3))
        loc = ops.linear(loc8, w8, self.image_location_embeddings.bias)
        rows = torch.arange(B * R, device=image_feature.device, dtype=torch.int32)
        y = ops.compose_ln(B * R, Hv, [(img, rows), (loc, rows)], [], self.LayerNorm.weight, self.LayerNorm.bias,
                           self.LayerNorm.eps, float(self.dropout.p), self.training)
        return y.view(B, R, Hv)


class B200BertTextEmbeddings(_BertEmbeddingsHolder):
    """HF BertEmbeddings (word + position + type -> LN -> dropout) through the composer; vilbert.py:1018"""

    def forward(self, input_ids, token_type_ids=None, position_ids=None):
        _require_cuda(input_ids, "input_ids")
        B, T = input_ids.shape
        dev = input_ids.device
        if position_ids is None:
            position_ids = torch.arange(T, device=dev).unsqueeze(0)
        if token_type_ids is None:
            token_type_ids = torch.zeros_like(input_ids)
        # HF BertEmbeddings broadcasts [1, T] ids over the batch (UNITER passes arange(T).unsqueeze(0), uniter.py:732-737):
        # the composer wants one index per output row
        position_ids = position_ids.expand(B, T)
        token_type_ids = token_type_ids.expand(B, T)
        H = self.LayerNorm.weight.shape[0]
        y = ops.compose_ln(B * T, H, [], [(self.word_embeddings.weight, ops.i32(input_ids), self.word_embeddings.padding_idx),
                                          (self.position_embeddings.weight, ops.i32(position_ids)),
                                          (self.token_type_embeddings.weight, ops.i32(token_type_ids))],
                           self.LayerNorm.weight, self.LayerNorm.bias, self.LayerNorm.eps, float(self.dropout.p),
                           self.training)
        return y.view(B, T, H)


class _FirstTokenPooler(nn.Module):
    """relu(dense(h[:, 0])): GEMM with the ReLU epilogue on the first-token rows"""

    def __init__(self, in_features, out_features):
        super().__init__()
        self.dense = nn.Linear(in_features, out_features)
        self.activation = nn.ReLU()

    def forward(self, hidden_states):
        return ops.linear_relu(hidden_states[:, 0], self.dense.weight, self.dense.bias).to(self.dense.weight.dtype)


class B200ViLBERTBase(nn.Module):
    """ViLBERTBase.forward (vilbert.py:936-1051) up to the encoder outputs; poolers / heads are torch consumers."""

    def __init__(self, config):
        super().__init__()
        self.config = config
        self.embeddings = B200BertTextEmbeddings(config)
        self.v_embeddings = B200ImageFeatureEmbeddings(config)
        self.encoder = B200ViLBertEncoder(config)
        self.t_pooler = _FirstTokenPooler(config.hidden_size, config.bi_hidden_size)        # BertTextPooler  vilbert.py:798-811
        self.v_pooler = _FirstTokenPooler(config.v_hidden_size, config.bi_hidden_size)      # BertImagePooler vilbert.py:814-827
        for m in (self.embeddings, self.v_embeddings, self.t_pooler, self.v_pooler):
            _init_bert_weights(m, float(getattr(config, "initializer_range", 0.02)))

    def forward(self, input_txt, image_feature, image_location, token_type_ids=None, attention_mask=None,
                image_attention_mask=None, co_attention_mask=None, task_ids=None, output_all_encoded_layers=False,
                output_all_attention_masks=False, reference_outputs=False):
        """-> (sequence_output_t, sequence_output_v, attention maps) by default; with reference_outputs=True the
        reference's 7-tuple (vilbert.py:1035-1051): (sequence_output_t, sequence_output_v, pooled_output_t,
        pooled_output_v, all_attention_mask, encoded_layers_t, encoded_layers_v) with the poolers applied."""
        if getattr(self.config, "task_specific_tokens", False) or task_ids is not None:
            # the reference prepends a mask column and adds a task embedding (vilbert.py:966-969): not built here
            raise NotImplementedError("ViLBERT task_specific_tokens / task_ids are not implemented on the B200 path")
        if output_all_encoded_layers and reference_outputs:
            # the reference's sequence_output_t/v would then be the state after the LAST CO-ATTENTION block, not the final
            # state (vilbert.py:1030-1033 with the encoder's per-block lists): refuse instead of silently differing
            raise NotImplementedError("output_all_encoded_layers=True with reference_outputs is not implemented")
        if attention_mask is None:
            attention_mask = torch.ones_like(input_txt)
        if token_type_ids is None:
            token_type_ids = torch.zeros_like(input_txt)
        if image_attention_mask is None:
            image_attention_mask = torch.ones(image_feature.size(0), image_feature.size(1), dtype=input_txt.dtype, device=input_txt.device)
        dt = self.embeddings.LayerNorm.weight.dtype
        ext_t = (1.0 - attention_mask.unsqueeze(1).unsqueeze(2).to(dt)) * -10000.0          # vilbert.py:982-1003
        ext_v = (1.0 - image_attention_mask.unsqueeze(1).unsqueeze(2).to(dt)) * -10000.0
        emb = self.embeddings(input_txt, token_type_ids)
        v_emb = self.v_embeddings(image_feature, image_location)
        t_layers, v_layers, attn = self.encoder(emb, v_emb, ext_t, ext_t, ext_v, co_attention_mask,
                                                output_all_encoded_layers=output_all_encoded_layers,
                                                output_all_attention_masks=output_all_attention_masks)
        seq_t, seq_v = t_layers[-1].to(dt), v_layers[-1].to(dt)
        if reference_outputs:
            return (seq_t, seq_v, self.t_pooler(seq_t), self.v_pooler(seq_v), attn if output_all_attention_masks else None,
                    t_layers if output_all_encoded_layers else None, v_layers if output_all_encoded_layers else None)
        return seq_t, seq_v, attn
