"""MMBT trunk on the B200 engine (BASELINE.json config 1).

  B200MMBTModel  <->  MMBTModel + ModalEmbeddings     mmf/models/mmbt.py:67-324
  B200MMBTBase   <->  MMBTBase.forward / extract_modal_end_token   mmf/models/mmbt.py:349-444

Sequence layout (mmbt.py:95-107, 227): [CLS-emb, projected regions, SEP-emb, text tokens shifted left by one]; the
modal rows and the text rows share the transformer's word/position/type tables and LayerNorm, so the whole
[B, L+T, H] embedding is ONE composer launch + one LayerNorm launch.  Integer token surgery is torch integer ops
(bit-exact with the reference, tests/golden/mmbt.pt).
"""
import torch
from torch import nn

from . import ops
from .modules import B200BertEncoder, _init_bert_weights, _require_cuda
from .visual_bert import BertPooler


class _BertEmbeddingsHolder(nn.Module):
    """parameter holder with HF BertEmbeddings' names (word/position/token_type embeddings, LayerNorm, dropout)"""

    def __init__(self, config):
        super().__init__()
        H = config.hidden_size
        # HF BertEmbeddings: padding_idx = pad_token_id, i.e. the [PAD] row gets no gradient
        self.word_embeddings = nn.Embedding(config.vocab_size, H, padding_idx=getattr(config, "pad_token_id", 0))
        self.position_embeddings = nn.Embedding(config.max_position_embeddings, H)
        self.token_type_embeddings = nn.Embedding(config.type_vocab_size, H)
        self.LayerNorm = nn.LayerNorm(H, eps=float(getattr(config, "layer_norm_eps", 1e-12)))
        self.dropout = nn.Dropout(float(config.hidden_dropout_prob))


class _BertModelHolder(nn.Module):
    """`transformer` of MMBTModel: embeddings + encoder + pooler (BertModelJit's children, hf_layers.py:371-375)"""

    def __init__(self, config):
        super().__init__()
        self.embeddings = _BertEmbeddingsHolder(config)
        self.encoder = B200BertEncoder(config)
        self.pooler = BertPooler(config.hidden_size)


class _ModalEmbeddingsHolder(nn.Module):
    """ModalEmbeddings (mmbt.py:67-83): its own projection + SHARED references to the transformer's tables"""

    def __init__(self, config, encoder, embeddings):
        super().__init__()
        self.encoder = encoder
        self.proj_embeddings = nn.Linear(config.modal_hidden_size, config.hidden_size)
        self.position_embeddings = embeddings.position_embeddings
        self.token_type_embeddings = embeddings.token_type_embeddings
        self.word_embeddings = embeddings.word_embeddings
        self.LayerNorm = embeddings.LayerNorm
        self.dropout = nn.Dropout(p=float(config.hidden_dropout_prob))


class B200MMBTModel(nn.Module):
    def __init__(self, config, modal_encoder=None, transformer=None):
        """`transformer`: an already built text encoder with children embeddings / encoder / pooler (what
        MMBTModel(config, transformer, encoder) receives, mmbt.py:148-156); built here when None."""
        super().__init__()
        self.config = config
        self.is_decoder = getattr(config, "is_decoder", False)
        if self.is_decoder:
            raise NotImplementedError("decoder (causal) MMBT is not on the B200 path")
        self.num_hidden_layers = config.num_hidden_layers
        fresh = transformer is None
        self.transformer = _BertModelHolder(config) if fresh else transformer
        self.modal_encoder = _ModalEmbeddingsHolder(config, modal_encoder or nn.Identity(), self.transformer.embeddings)
        std = float(getattr(config, "initializer_range", 0.02))
        if fresh:
            _init_bert_weights(self.transformer, std)
        _init_bert_weights(self.modal_encoder.proj_embeddings, std)     # passed-in encoders keep their weights

    def forward(self, input_modal, input_ids, modal_start_tokens=None, modal_end_tokens=None, attention_mask=None,
                token_type_ids=None, modal_token_type_ids=None, position_ids=None, modal_position_ids=None,
                head_mask=None, inputs_embeds=None, encoder_hidden_states=None, encoder_attention_mask=None):
        """-> (sequence_output, pooled_output, ())   same argument meaning as MMBTModel.forward (mmbt.py:176-318)"""
        if inputs_embeds is not None or head_mask is not None or encoder_hidden_states is not None:
            raise NotImplementedError("inputs_embeds / head_mask / encoder_hidden_states are not on the B200 path")
        if input_ids is None:
            raise ValueError("You have to specify either input_ids or inputs_embeds")
        _require_cuda(input_ids, "input_ids")
        emb = self.transformer.embeddings
        dev = input_ids.device
        B, T = input_ids.shape
        modal = self.modal_encoder.encoder(input_modal)
        if modal.dim() == 2:
            modal = modal.unsqueeze(1)
        R, Fdim = modal.shape[1], modal.shape[2]
        H = self.config.hidden_size
        proj = ops.linear(modal.reshape(B * R, Fdim), self.modal_encoder.proj_embeddings.weight,
                          self.modal_encoder.proj_embeddings.bias)
        hs, he = int(modal_start_tokens is not None), int(modal_end_tokens is not None)
        L = R + hs + he
        # ---- integer index construction (exact) ----
        neg = lambda *s: torch.full(s, -1, dtype=torch.long, device=dev)
        word_m = neg(B, L)
        if hs:
            word_m[:, 0] = modal_start_tokens
        if he:
            word_m[:, L - 1] = modal_end_tokens
        src_m = neg(B, L)
        src_m[:, hs:hs + R] = torch.arange(B * R, device=dev).view(B, R)
        pos_m = (modal_position_ids if modal_position_ids is not None
                 else torch.arange(L, device=dev).unsqueeze(0).expand(B, L))
        if modal_token_type_ids is None:
            modal_token_type_ids = torch.zeros((B, L), dtype=torch.long, device=dev)
        type_m = modal_token_type_ids.expand(B, L) if modal_token_type_ids.shape[1] == 1 else modal_token_type_ids
        if token_type_ids is None:
            token_type_ids = torch.ones((B, T), dtype=torch.long, device=dev)    # mmbt.py:215-218
        pos_t = position_ids if position_ids is not None else torch.arange(T, device=dev).unsqueeze(0).expand(B, T)
        word = torch.cat([word_m, input_ids], dim=1)
        src = torch.cat([src_m, neg(B, T)], dim=1)
        pos = torch.cat([pos_m, pos_t], dim=1)
        typ = torch.cat([type_m, token_type_ids], dim=1)
        S = L + T
        x = ops.compose_ln(B * S, H, srcs=[(proj, ops.i32(src))],
                           tabs=[(emb.word_embeddings.weight, ops.i32(word), emb.word_embeddings.padding_idx), (emb.position_embeddings.weight, ops.i32(pos)),
                                 (emb.token_type_embeddings.weight, ops.i32(typ))],
                           ln_weight=emb.LayerNorm.weight, ln_bias=emb.LayerNorm.bias, eps=emb.LayerNorm.eps,
                           p=float(emb.dropout.p), training=self.training).view(B, S, H)
        # ---- attention mask (mmbt.py:231-285) ----
        if attention_mask is None:
            full = torch.ones((B, S), dtype=torch.long, device=dev)
        else:
            full = torch.cat([torch.ones((B, L), device=dev, dtype=torch.long), attention_mask], dim=1)
        ext = (1.0 - full[:, None, None, :].to(emb.LayerNorm.weight.dtype)) * -10000.0
        seq = self.transformer.encoder(x, ext)[0].to(emb.LayerNorm.weight.dtype)
        return seq, self.transformer.pooler(seq), ()


def extract_modal_end_token(sample_list):
    """Semantics of MMBTBase.extract_modal_end_token (mmbt.py:349-374): per sample, the id of the last unmasked
    token (the [SEP]) is returned, and the text is shifted left by one position in place - the leading [CLS] moves to
    the modal block, the last id is repeated, the mask gains a trailing 0.  Pure integer indexing (bit-exact)."""
    ids, mask = sample_list["input_ids"], sample_list["input_mask"]
    last = mask.sum(dim=1) - 1                                   # index of the last attended token
    end_token = ids[torch.arange(ids.size(0), device=ids.device), last].clone()
    sample_list["input_ids"] = torch.cat([ids[:, 1:], ids[:, -1:]], dim=1)     # shift left, the last id repeated
    shifted_mask = torch.roll(mask, shifts=-1, dims=1)
    shifted_mask[:, -1] = 0
    sample_list["input_mask"] = shifted_mask
    return end_token


class B200MMBTBase(nn.Module):
    """MMBTBase.forward (mmbt.py:376-444) for direct feature input."""

    def __init__(self, config, use_modal_start_token=True, use_modal_end_token=True, num_max_segment=2,
                 modal_encoder=None, transformer=None):
        super().__init__()
        self.mmbt = B200MMBTModel(config, modal_encoder=modal_encoder, transformer=transformer)
        self.use_modal_start_token, self.use_modal_end_token = use_modal_start_token, use_modal_end_token
        self.num_max_segment = num_max_segment

    @classmethod
    def from_config(cls, config):
        """The reference's construction route (MMBTBase.build, mmbt.py:333-347): `text_encoder` / `modal_encoder`
        factory configs -> encoders (mmf_b200.encoders) -> MMBTModel(mmbt_config, text_encoder, modal_encoder).
        Only `direct_features_input: true` is on the fusion path (e.g. configs/models/mmbt/with_features.yaml)."""
        import types
        from .encoders import B200MultiModalEncoderBase, _get
        enc = B200MultiModalEncoderBase(config)
        if enc.text_encoder is None or not hasattr(enc.text_encoder, "config"):
            raise ValueError("MMBT needs a transformer text_encoder")
        mm = types.SimpleNamespace(**vars(enc.encoder_config))
        mm.modal_hidden_size = _get(config, "modal_hidden_size", 2048)      # MMBTConfig (mmbt.py:338-342)
        mm.num_labels = _get(config, "num_labels", None)
        te_params = _get(_get(config, "text_encoder"), "params")
        return cls(mm, _get(config, "use_modal_start_token", True), _get(config, "use_modal_end_token", True),
                   _get(te_params, "num_segments", None) or 2, modal_encoder=enc.modal_encoder,
                   transformer=enc.text_encoder)

    def forward(self, sample_list):
        input_modal = sample_list["input_modal"] if "input_modal" in sample_list else sample_list["image_feature_0"]
        start = sample_list["input_ids"][:, 0].clone().detach() if self.use_modal_start_token else None
        end = extract_modal_end_token(sample_list) if self.use_modal_end_token else None
        if "modal_token_type_ids" in sample_list:
            modal_tt = sample_list["modal_token_type_ids"]
        else:
            # segment id given to the modal block (mmbt.py:393-414): with a single text segment it is "the other one"
            # (1 if the text uses 0, else 0); with several, the last segment unless the text already ends there
            # computed on the device (no host round trip in the step): the reference's Python branches on
            # int(seg.min()) / int(seg.max()) as tensor selects
            seg = sample_list["segment_ids"]
            lo, hi = seg.min(), seg.max()
            top = self.num_max_segment - 1
            single = torch.where(hi == 0, torch.ones_like(hi), torch.zeros_like(hi))
            multi = torch.where(hi != top, torch.full_like(hi, top), torch.zeros_like(hi))
            token_value = torch.where(lo == hi, single, multi).to(torch.long)
            modal_tt = token_value.reshape(1, 1).expand(input_modal.size(0), 1).to(input_modal.device)
        if input_modal.dim() == 2:
            input_modal = input_modal.unsqueeze(dim=1)
        return self.mmbt(input_modal, input_ids=sample_list["input_ids"], modal_start_tokens=start,
                         modal_end_tokens=end, attention_mask=sample_list["input_mask"],
                         token_type_ids=sample_list["segment_ids"], modal_token_type_ids=modal_tt)
