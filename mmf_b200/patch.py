"""Layer-implementation swap by monkey-patch - the reference's own mechanism for this boundary:
`replace_with_jit()` / `undo_replace_with_jit()` (mmf/modules/hf_layers.py:48-93, mmf/utils/patch.py:198-244) rebind
`BertEncoder.forward` & co process-wide from model constructors.  `replace_with_b200()` rebinds the forward of
HuggingFace `BertEncoder` (and every subclass that overrides `forward`, e.g. MMF's `BertEncoderJit`) to the B200 engine
with the contract of `BertEncoderJit.forward` (hf_layers.py:316-355).  Parameters keep their names; their storage is
re-homed into the engine's flat pack on the first CUDA forward.
"""
import types

from .modules import EncoderRunner, run_bert_encoder

_saved = {}


def _b200_encoder_forward(self, hidden_states, attention_mask=None, encoder_hidden_states=None,
                          encoder_attention_mask=None, output_attentions=False, output_hidden_states=False,
                          return_dict=False, head_mask=None, **unused):
    if output_attentions:
        raise NotImplementedError("output_attentions: the fused kernel never materialises attention probabilities")
    if head_mask is not None and not (isinstance(head_mask, (list, tuple)) and all(h is None for h in head_mask)):
        raise NotImplementedError("head_mask is not supported on the B200 path")
    runner = _runner_of(self)
    outs = run_bert_encoder(runner, hidden_states, attention_mask, self.training, output_hidden_states)
    if output_hidden_states:
        return (outs[0], tuple(outs[1:]) + (outs[0],))
    return (outs[0],)


def _runner_of(encoder):
    r = encoder.__dict__.get("_runner")
    if r is None:
        r = EncoderRunner(encoder.layer)
        object.__setattr__(encoder, "_runner", r)
    return r


def attach_encoder(encoder):
    """Route ONE encoder instance (anything with `.layer[i].attention.self.{query,key,value}` ...) through the engine."""
    _runner_of(encoder)
    encoder.forward = types.MethodType(_b200_encoder_forward, encoder)
    return encoder


def _all_subclasses(cls):
    out = []
    for sub in cls.__subclasses__():
        out.append(sub)
        out += _all_subclasses(sub)
    return out


def replace_with_b200():
    """Process-wide swap, mirroring replace_with_jit()."""
    from transformers.models.bert.modeling_bert import BertEncoder
    for cls in [BertEncoder] + _all_subclasses(BertEncoder):
        if "forward" in cls.__dict__ and cls not in _saved:
            _saved[cls] = cls.__dict__["forward"]
            cls.forward = _b200_encoder_forward


def undo_replace_with_b200():
    """mirrors undo_replace_with_jit()"""
    for cls, fwd in list(_saved.items()):
        cls.forward = fwd
        del _saved[cls]
