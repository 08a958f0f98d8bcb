"""MMFTransformer model plugin on the B200 engine (BASELINE.json configs[3]).

  registry.register_model("mmft") / ("mmf_transformer")  MMFTransformer  <->  mmf/models/mmf_transformer.py:35-442
      preprocess_sample + the `_infer_*` family (:176-392): per-modality input ids / position ids / segment ids / masks /
      MLM labels (with `combined_labels`) / ITM labels - INTEGER paths, bit-exact with the reference's known answers
      (its tests/models/test_mmf_transformer.py:205-402, mirrored in tests/test_mmft_cpu.py)
      forward (:394-419): backend(input_ids, position_ids, segment_ids, masks) -> heads
  heads (registry "transformer_head" table of the reference):
      "mlp"  MLPHead  <-> mmf/models/transformers/heads/mlp.py:19-96    pooler -> [dropout, transform] x n -> Linear
      "mlm"  MLMHead  <-> heads/mlm.py:20-97     labelled positions gathered FIRST, transform -> vocabulary decoder -> CE
      "itm"  ITMHead  <-> heads/itm.py:19-74     pooler -> Linear(H, 2) -> CE

The backend is the registered "b200" transformer backend (mmf_b200.mmft_backend); a config that says `backend.type:
huggingface` (the reference default, configs/models/mmf_transformer/defaults.yaml:4-7) is served by it as well - that is
the drop-in.  `transformer_base` cannot be downloaded here: the BERT-base hyper-parameters are used and overridden by
`backend.params` / `transformer_config` keys; weights arrive through load_state_dict.
"""
import warnings

import torch
from torch import nn

from . import ops
from .encoders import build_encoder
from .heads import B200BertLMPredictionHead, B200BertPredictionHeadTransform
from .models import BERT_BASE, MODEL_DEFAULTS, B200Linear, BaseModel, ConfigNode
from .modules import _init_bert_weights
from .registry import registry
from .visual_bert import BertPooler

# mmf/configs/models/mmf_transformer/defaults.yaml:1-41 (the heads' num_labels interpolation resolved)
MODEL_DEFAULTS["mmf_transformer"] = dict(
    transformer_base="bert-base-uncased", backend=dict(type="huggingface", freeze=False, params={}),
    heads=[dict(type="mlp", freeze=False, lr_multiplier=1.0, hidden_size=768, num_labels=2)],
    modalities=[
        dict(type="text", key="text", position_dim=512, segment_id=0, embedding_dim=768, layer_norm_eps=1e-12,
             hidden_dropout_prob=0.1),
        dict(type="image", key="image", embedding_dim=2048, position_dim=1, segment_id=1, layer_norm_eps=1e-12,
             hidden_dropout_prob=0.1,
             encoder=dict(type="resnet152", params=dict(pretrained=True, pool_type="avg", num_output_features=1)))],
    initializer_range=0.02, initializer_mean=0.0, token_noise_std=0.01, token_noise_mean=0.0, layer_norm_weight_fill=1.0,
    random_initialize=False, freeze_image_encoder=False, tie_weight_to_encoder=None, num_labels=2)
MODEL_DEFAULTS["mmft"] = MODEL_DEFAULTS["mmf_transformer"]

HEADS = {}


def register_transformer_head(name):
    def wrap(cls):
        HEADS[name] = cls
        return cls
    return wrap


# ------------------------------------------------------------------------------------------------------
# heads
# ------------------------------------------------------------------------------------------------------
class _TransformWithInDim(B200BertPredictionHeadTransform):
    """PredictionHeadTransformWithInDim (mlp.py:90-93): the transform's dense maps in_dim -> hidden_size"""

    def __init__(self, config):
        super().__init__(config)
        self.dense = nn.Linear(config.in_dim, config.hidden_size)


@register_transformer_head("multilayer_mlp")
@register_transformer_head("mlp")
class MLPHead(nn.Module):
    DEFAULTS = dict(type="mlp", num_labels=2, hidden_size=768, hidden_dropout_prob=0.1, layer_norm_eps=1e-6,
                    hidden_act="gelu", pooler_name="bert_pooler", num_layers=1, in_dim=None)      # mlp.py:22-32

    def __init__(self, config):
        super().__init__()
        self.config = c = ConfigNode(dict(self.DEFAULTS, **dict(config)))
        self.num_labels, self.hidden_size = c.num_labels, c.hidden_size
        self.in_dim = c.in_dim = c.hidden_size if c.in_dim is None else c.in_dim
        if c.pooler_name == "bert_pooler":
            self.pooler = BertPooler(self.in_dim)
        elif c.pooler_name == "identity":
            self.pooler = nn.Identity()
        else:
            raise NotImplementedError("%s is not implemented." % c.pooler_name)
        if c.num_layers < 0:
            raise AssertionError("num_layers must be >= 0")
        layers, hc = [], ConfigNode(dict(c))
        for _ in range(c.num_layers):
            layers.append(nn.Dropout(c.hidden_dropout_prob))
            layers.append(_TransformWithInDim(ConfigNode(dict(hc))))
            hc.in_dim = hc.hidden_size
        self.classifier = nn.Sequential(*layers, B200Linear(self.hidden_size, self.num_labels))

    def forward(self, sequence_output, encoded_layers=None, processed_sample_list=None):
        assert sequence_output.size(-1) == self.in_dim, "Mismatch between MLP head hidden_size and sequence_output last dim."
        pooled = self.pooler(sequence_output)
        return {"scores": self.classifier(pooled).reshape(-1, self.num_labels).to(sequence_output.dtype)}


@register_transformer_head("mlm")
class MLMHead(nn.Module):
    DEFAULTS = dict(type="mlm", vocab_size=30522, hidden_size=768, hidden_dropout_prob=0.1, layer_norm_eps=1e-5,
                    hidden_act="gelu", ignore_index=-1, loss_name="masked_lm_loss", label_key=None)   # mlm.py:23-32

    def __init__(self, config):
        super().__init__()
        self.config = c = ConfigNode(dict(self.DEFAULTS, **dict(config)))
        self.cls = nn.Module()                                     # BertOnlyMLMHead: `cls.predictions.*`
        self.cls.predictions = B200BertLMPredictionHead(c)
        self.vocab_size = c.vocab_size

    def tie_weights(self, module=None):
        self.cls.predictions.decoder.weight = module.weight

    def forward(self, sequence_output, encoded_layers=None, processed_sample_list=None):
        assert processed_sample_list is not None, "MLM head requires 'processed_sample_list' argument"
        c = self.config
        if c.label_key is not None:
            assert c.label_key in processed_sample_list
            labels = processed_sample_list[c.label_key]
        else:
            assert processed_sample_list.get("mlm_labels") is not None
            assert "combined_labels" in processed_sample_list["mlm_labels"]
            labels = processed_sample_list["mlm_labels"]["combined_labels"]
        sel = labels.ne(c.ignore_index)                             # gather the labelled rows first (mlm.py:79-82)
        labels = labels[sel]
        if labels.numel() == 0:
            # CrossEntropyLoss over zero targets is NaN, which the reference replaces by 0 (mlm.py:89-93)
            warnings.warn("NaN detected in masked_lm_loss. Replacing it with 0.")
            return {"logits": sequence_output.new_zeros(0, self.vocab_size),
                    "losses": {c.loss_name: sequence_output.sum() * 0.0}}
        rows = sequence_output[sel, :]
        logits = self.cls.predictions(rows)
        loss = nn.functional.cross_entropy(logits.reshape(-1, self.vocab_size).float(), labels.reshape(-1),
                                           ignore_index=c.ignore_index)
        return {"logits": logits, "losses": {c.loss_name: loss}}


@register_transformer_head("itm")
class ITMHead(nn.Module):
    DEFAULTS = dict(type="itm", hidden_size=768, loss_name="itm_loss", ignore_index=-1, itm_label_key="is_correct")   # itm.py:22-27

    def __init__(self, config):
        super().__init__()
        self.config = c = ConfigNode(dict(self.DEFAULTS, **dict(config)))
        self.pooler = BertPooler(c.hidden_size)
        self.cls = nn.Module()                                     # BertOnlyNSPHead: `cls.seq_relationship`
        self.cls.seq_relationship = B200Linear(c.hidden_size, 2)

    def forward(self, sequence_output, encoded_layers=None, processed_sample_list=None):
        assert processed_sample_list is not None, "ITM head requires 'processed_sample_list' argument"
        c = self.config
        if c.itm_label_key in processed_sample_list:
            labels = processed_sample_list[c.itm_label_key]
        else:
            assert processed_sample_list.get("itm_labels") is not None
            labels = processed_sample_list["itm_labels"][c.itm_label_key]
        score = self.cls.seq_relationship(self.pooler(sequence_output))
        loss = nn.functional.cross_entropy(score.reshape(-1, 2).float(), labels.reshape(-1), ignore_index=c.ignore_index)
        return {"losses": {c.loss_name: loss}}


# ------------------------------------------------------------------------------------------------------
# the model
# ------------------------------------------------------------------------------------------------------
def _text_slice(tensor, text_index):
    """stacked text fields [B, n_text, L] hold one row per text modality, in config order (mmf_transformer.py:233-241)"""
    return (tensor[:, text_index], text_index + 1) if tensor.dim() > 2 else (tensor, text_index)


@registry.register_model("mmft")
@registry.register_model("mmf_transformer")
class MMFTransformer(BaseModel):
    def __init__(self, config, *args, **kwargs):
        super().__init__(config)
        self.modality_keys, self.modality_type, self.modality_segments = [], [], []
        for m in self.config.modalities:
            self.modality_keys.append(m["key"])
            self.modality_type.append(m["type"])
            self.modality_segments.append(m["segment_id"] if "segment_id" in m else -1)

    @classmethod
    def config_path(cls):
        return "configs/models/mmf_transformer/defaults.yaml"

    @classmethod
    def format_state_key(cls, key):                                # mmf_transformer.py:97-104
        if key.startswith("pooler.") or key.startswith("classifier."):
            return key.replace("pooler.", "heads.0.pooler.").replace("classifier.", "heads.0.classifier.")
        return key

    # ---- construction (transformers/base.py:100-111, 170-238) ----
    def build(self):
        self.build_backend()
        self.build_encoders()
        self.build_heads()
        self.init_weights()

    def _transformer_config(self):
        tc = dict(BERT_BASE)
        tc.update(dict(self.config.get("transformer_config", {}) or {}))
        tc.update(dict((self.config.get("backend", {}) or {}).get("params", {}) or {}))
        return ConfigNode(tc)

    def build_backend(self):
        backend_config = self.config.get("backend", {}) or {}
        btype = backend_config.get("type", "huggingface")
        # the reference's default type is served by the B200 backend: that IS the drop-in for this path
        from . import mmft_backend  # noqa: F401  (registers "b200")
        cls = registry.get_transformer_backend_class("b200" if btype == "huggingface" else btype)
        if cls is None:
            raise RuntimeError("no transformer backend registered under %r" % btype)
        bc = ConfigNode(dict(self.config))
        bc.transformer_config = self._transformer_config()
        self.backend = cls(bc)
        if backend_config.get("freeze", False):
            for p in self.backend.parameters():
                p.requires_grad = False

    def build_encoders(self):
        self.encoders = nn.ModuleDict()
        for m in self.config.modalities:
            if "encoder" not in m or m["encoder"] is None:
                if m["type"] == "image" and "image_encoder" in self.config:
                    ec = self.config.image_encoder
                else:
                    ec = {"type": "identity", "params": {"in_dim": 100}}
            else:
                ec = m["encoder"]
            enc = build_encoder(ec)
            self.encoders[m["key"]] = enc
            frozen = (m["type"] == "image" and self.config.get("freeze_image_encoder", False)) or (
                m["type"] == "text" and self.config.get("freeze_text_encoder", False))
            if frozen:
                for p in enc.parameters():
                    p.requires_grad = False

    def build_heads(self):
        self.heads = nn.ModuleList()
        for hc in self.config.get("heads", []):
            htype = hc.get("type", "mlp")
            if htype not in HEADS:
                raise NotImplementedError("transformer head %r is not on the B200 path (have: %s)" % (htype, sorted(HEADS)))
            self.heads.append(HEADS[htype](hc))

    def init_weights(self):
        if self.config.get("random_initialize", False) is False and self.config.get("transformer_base", None) is None:
            _init_bert_weights(self.heads, float(self.config.get("initializer_range", 0.02)))
        self.tie_weights()

    def tie_weights(self):
        """mmf_transformer.py:145-174: heads with a tie_weights() share the text modality's token table"""
        if "text" not in self.modality_type:
            return
        idx = self.modality_type.index("text")
        for head in self.heads:
            if hasattr(head, "tie_weights"):
                head.tie_weights(self.backend.embeddings.token_embeddings[idx])

    # ---- SampleList -> per-modality integer tensors (mmf_transformer.py:176-392) ----
    @staticmethod
    def _first_present(sample_list, keys):
        for k in keys:
            if k in sample_list:
                return sample_list[k]
        expected = keys[0] if len(keys) == 1 else "%s or %s" % (", ".join(keys[:-1]), keys[-1])
        raise TypeError("Missing modality in SampleList. Expected to find %s" % expected)

    def _infer_input_ids(self, sample_list):
        out, ti = {}, 0
        for idx, (key, enc) in enumerate(zip(self.modality_keys, self.encoders.values())):
            kind = self.modality_type[idx]
            if kind == "text":
                out[key], ti = _text_slice(self._first_present(sample_list, ("input_ids", key)), ti)
            elif kind == "image":
                out[key] = self._first_present(sample_list, (key, "image", "input_modal", "image_feature_0"))
            else:
                out[key] = self._first_present(sample_list, (key,))
            if enc is not None:
                out[key] = enc(out[key])
            if kind != "text" and out[key].dim() == 2:             # [B, D] feature = one position
                out[key] = out[key].unsqueeze(1)
        return out

    def _infer_position_ids(self, input_ids):
        out = {}
        for key in self.modality_keys:
            B, n = input_ids[key].size(0), input_ids[key].size(1)
            out[key] = torch.arange(0, n, dtype=torch.long, device=input_ids[key].device).unsqueeze(0).expand((B, n))
        return out

    def _infer_masks(self, sample_list, input_ids):
        out, ti = {}, 0
        for idx, key in enumerate(self.modality_keys):
            if self.modality_type[idx] == "text" and "input_mask" in sample_list:
                out[key], ti = _text_slice(sample_list["input_mask"], ti)
            elif (key + "_mask") in sample_list:
                out[key] = sample_list[key + "_mask"]
            else:
                out[key] = torch.ones(input_ids[key].size()[:2], dtype=torch.long, device=input_ids[key].device)
        return out

    def _infer_segment_ids(self, sample_list, input_ids):
        out, ti = {}, 0
        for idx, key in enumerate(self.modality_keys):
            if self.modality_segments[idx] == -1:
                continue
            if self.modality_type[idx] == "text" and "segment_ids" in sample_list:
                out[key], ti = _text_slice(sample_list["segment_ids"], ti)
            else:
                out[key] = torch.full(input_ids[key].size()[:2], fill_value=self.modality_segments[idx], dtype=torch.long,
                                      device=input_ids[key].device)
        return out

    def _infer_itm_labels(self, sample_list, input_ids):
        if "is_correct" in sample_list:
            return {"is_correct": sample_list["is_correct"]}
        return {"is_correct": torch.tensor(True, dtype=torch.long, device=input_ids[self.modality_keys[0]].device)}

    def _infer_mlm_labels(self, sample_list, input_ids):
        out, ti = {}, 0
        for idx, key in enumerate(self.modality_keys):
            if self.modality_type[idx] == "text" and "lm_label_ids" in sample_list:
                out[key], ti = _text_slice(sample_list["lm_label_ids"], ti)
            else:
                out[key] = torch.full(input_ids[key].size()[:2], fill_value=-1, dtype=torch.long, device=input_ids[key].device)
        if self.modality_keys:
            out["combined_labels"] = torch.cat([out[k] for k in self.modality_keys], dim=-1)
        return out

    def preprocess_sample(self, sample_list):
        input_ids = self._infer_input_ids(sample_list)
        return {"input_ids": input_ids, "position_ids": self._infer_position_ids(input_ids),
                "segment_ids": self._infer_segment_ids(sample_list, input_ids),
                "masks": self._infer_masks(sample_list, input_ids),
                "mlm_labels": self._infer_mlm_labels(sample_list, input_ids),
                "itm_labels": self._infer_itm_labels(sample_list, input_ids)}

    def forward(self, sample_list):
        processed = self.preprocess_sample(sample_list)
        processed["target_key"] = sample_list
        masks = [processed["masks"][k] for k in self.modality_keys]
        sequence_output, encoded_layers = self.backend(processed["input_ids"], processed["position_ids"],
                                                       processed["segment_ids"], masks)
        return self.postprocess_output(sequence_output, encoded_layers, processed)

    def postprocess_output(self, sequence_output, encoded_layers, processed_sample_list):
        out = {}
        for head in self.heads:
            out.update(head(sequence_output, encoded_layers, processed_sample_list))
        return out
