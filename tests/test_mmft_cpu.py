"""MMFTransformer.preprocess_sample - the integer id / position / segment / mask / label inference of
mmf/models/mmf_transformer.py:176-392 - against the KNOWN ANSWERS of the reference's own tests
(/root/reference/tests/models/test_mmf_transformer.py:205-402: the same modality configs, the same SampleLists, the same
expected tensors, compared with torch.equal), plus the head wiring over the kernel test double."""
import os
import sys

import pytest
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import fake_kernels as FK  # noqa: E402

SMALL = {"hidden_size": 64, "num_hidden_layers": 1, "num_attention_heads": 1, "intermediate_size": 128, "vocab_size": 512,
         "max_position_embeddings": 128, "hidden_dropout_prob": 0.0, "attention_probs_dropout_prob": 0.0}


def _image():   # test_mmf_transformer.py:180-187
    return dict(type="image", key="image", embedding_dim=256, position_dim=1, segment_id=0, encoder=dict(type="identity"))


def _text(**over):   # :188-195
    d = dict(type="text", key="text", embedding_dim=756, position_dim=128, segment_id=1, encoder=dict(type="identity"))
    d.update(over)
    return d


def _build(modalities, heads=None):
    from mmf_b200 import models as MD
    import mmf_b200.mmft  # noqa: F401
    cfg = MD.load_model_config("mmf_transformer", overrides={"modalities": modalities, "num_labels": 2,
                                                             "transformer_config": SMALL,
                                                             "heads": heads or [{"type": "mlp", "hidden_size": 64, "num_labels": 2}]})
    assert MD.registry.get_model_class("mmft") is MD.registry.get_model_class("mmf_transformer")
    return MD.build_model(cfg)


def _sl(**kw):
    from mmf_b200.sample import SampleList
    return SampleList(**kw)


def eq(a, b):
    assert a.dtype == b.dtype and torch.equal(a, b), (a, b)


def test_one_dim_feature_preprocessing():
    """reference test :205-242"""
    mmft = _build([_image(), _text()])
    sl = _sl(image=torch.rand(2, 256), text=torch.randint(0, 512, (2, 128)))
    t = mmft.preprocess_sample(sl)
    assert list(t["input_ids"]["image"].size()) == [2, 1, 256] and list(t["input_ids"]["text"].size()) == [2, 128]
    eq(t["position_ids"]["image"], torch.tensor([[0], [0]]))
    eq(t["position_ids"]["text"], torch.arange(0, 128).unsqueeze(0).expand((2, 128)))
    eq(t["masks"]["image"], torch.tensor([[1], [1]]))
    eq(t["masks"]["text"], torch.ones((2, 128)).long())
    eq(t["segment_ids"]["image"], torch.tensor([[0], [0]]))
    eq(t["segment_ids"]["text"], torch.ones((2, 128)).long())
    eq(t["mlm_labels"]["combined_labels"], torch.full((2, 129), dtype=torch.long, fill_value=-1))
    eq(t["itm_labels"]["is_correct"], torch.tensor(True, dtype=torch.long))


def _compare_multimodality(t, lm_labels_sum):
    """reference helper :296-338"""
    ids = t["input_ids"]
    assert list(ids["image"].size()) == [2, 1, 256] and list(ids["body"].size()) == [2, 128] and list(ids["ocr"].size()) == [2, 128]
    eq(t["position_ids"]["image"], torch.tensor([[0], [0]]))
    for k in ("body", "ocr"):
        eq(t["position_ids"][k], torch.arange(0, 128).unsqueeze(0).expand((2, 128)))
        eq(t["masks"][k], torch.ones((2, 128)).long())
    eq(t["masks"]["image"], torch.tensor([[1], [1]]))
    eq(t["segment_ids"]["image"], torch.tensor([[0], [0]]))
    eq(t["segment_ids"]["body"], torch.ones((2, 128)).long())
    eq(t["segment_ids"]["ocr"], torch.full((2, 128), dtype=torch.long, fill_value=2))
    assert list(t["mlm_labels"]["combined_labels"].size()) == [2, 257]
    assert t["mlm_labels"]["combined_labels"].sum().item() == lm_labels_sum - 2     # -2: the image's two -1 labels


def _three():
    return [_image(), _text(key="body"), _text(key="ocr", segment_id=2)]


def test_stacked_feature_preprocessing():
    """reference test :244-271: stacked [B, 2, L] text fields are split in modality order"""
    mmft = _build(_three())
    g = torch.Generator().manual_seed(0)
    ids = torch.randint(0, 512, (2, 2, 128), generator=g)
    lab = torch.randint(-1, 30522, (2, 2, 128), generator=g)
    t = mmft.preprocess_sample(_sl(image=torch.rand(2, 256), input_ids=ids, lm_label_ids=lab))
    _compare_multimodality(t, lab.sum().item())
    eq(t["input_ids"]["body"], ids[:, 0])
    eq(t["input_ids"]["ocr"], ids[:, 1])
    eq(t["mlm_labels"]["combined_labels"], torch.cat([torch.full((2, 1), -1), lab[:, 0], lab[:, 1]], dim=-1))


def test_modality_key_preprocessing():
    """reference test :273-294: per-key text fields, one shared 2-D lm_label_ids used for both"""
    mmft = _build(_three())
    lab = torch.randint(-1, 30522, (2, 128))
    t = mmft.preprocess_sample(_sl(image=torch.rand(2, 256), body=torch.randint(0, 512, (2, 128)),
                                   ocr=torch.randint(0, 512, (2, 128)), lm_label_ids=lab))
    _compare_multimodality(t, lab.sum().item() * 2)


def test_custom_feature_and_mask_preprocessing():
    """reference test :340-402"""
    extra = dict(type="my_random_feature", key="my_random_feature", embedding_dim=128, position_dim=4, segment_id=3,
                 encoder=dict(type="identity"))
    mmft = _build([_image(), _text(), extra])
    sl = _sl(image=torch.rand(2, 256), text=torch.randint(0, 512, (2, 128)), text_mask=torch.ones(2, 128),
             my_random_feature=torch.rand(2, 4, 128), my_random_feature_mask=torch.ones(2, 4))
    sl.text_mask[:, 70:] = 0
    sl.my_random_feature_mask[:, 3:] = 0
    t = mmft.preprocess_sample(sl)
    assert list(t["input_ids"]["my_random_feature"].size()) == [2, 4, 128]
    eq(t["position_ids"]["my_random_feature"], torch.arange(0, 4).unsqueeze(0).expand((2, 4)))
    eq(t["masks"]["image"], torch.tensor([[1], [1]]))
    assert t["masks"]["text"].sum().item() == 140 and t["masks"]["my_random_feature"].sum().item() == 6
    eq(t["segment_ids"]["text"], torch.ones((2, 128)).long())
    eq(t["segment_ids"]["my_random_feature"], torch.full((2, 4), dtype=torch.long, fill_value=3))


def test_missing_modality_raises_typeerror():
    """mmf_transformer.py:262-278"""
    mmft = _build([_image(), _text()])
    with pytest.raises(TypeError, match="Expected to find image, image, input_modal or image_feature_0"):
        mmft.preprocess_sample(_sl(text=torch.randint(0, 512, (2, 128))))


def test_forward_heads_over_the_test_double(monkeypatch):
    """backend -> mlp / mlm / itm heads end to end on the CPU double: scores shape, MLM evaluated on the labelled rows only
    (logits [n_labelled, V]), decoder tied to the text token table, losses under the heads' names."""
    import mmf_b200.embeddings as EM
    import mmf_b200.engine as E
    import mmf_b200.heads as HD
    import mmf_b200.mmft_backend as MF
    import mmf_b200.modules as M
    import mmf_b200.ops as OPS
    monkeypatch.setattr(E, "F", FK)
    monkeypatch.setattr(OPS, "F", FK)
    for mod in (M, EM, MF, HD):
        if hasattr(mod, "_require_cuda"):
            monkeypatch.setattr(mod, "_require_cuda", lambda t, what: None)
    mods = [dict(type="text", key="text", embedding_dim=64, position_dim=128, segment_id=0),
            dict(type="image", key="image", embedding_dim=48, position_dim=8, segment_id=1, encoder=dict(type="identity"))]
    heads = [{"type": "mlp", "hidden_size": 64, "num_labels": 3},
             {"type": "mlm", "hidden_size": 64, "vocab_size": 512},
             ]
    torch.manual_seed(3)
    m = _build(mods, heads).eval()
    assert m.heads[1].cls.predictions.decoder.weight is m.backend.embeddings.token_embeddings[0].weight    # tie_weights
    lab = torch.full((2, 10), -1)
    lab[0, 3], lab[1, 7], lab[1, 8] = 5, 9, 400
    sl = _sl(input_ids=torch.randint(1, 512, (2, 10)), input_mask=torch.ones(2, 10, dtype=torch.long),
             image=torch.rand(2, 6, 48), lm_label_ids=lab)
    out = m(sl)
    assert out["scores"].shape == (2, 3) and out["logits"].shape == (3, 512)
    assert list(out["losses"]) == ["masked_lm_loss"] and torch.isfinite(out["losses"]["masked_lm_loss"])
    out["losses"]["masked_lm_loss"].backward()
    assert m.backend.embeddings.token_embeddings[0].weight.grad is not None
    # no labelled position: zero loss, no GEMM with M = 0 (the reference replaces the NaN by 0, mlm.py:89-93)
    sl2 = _sl(input_ids=sl["input_ids"], input_mask=sl["input_mask"], image=sl["image"], lm_label_ids=torch.full((2, 10), -1))
    with pytest.warns(UserWarning):
        assert float(m(sl2)["losses"]["masked_lm_loss"]) == 0.0
