"""TEST / BASELINE INFRASTRUCTURE - loads the reference's own hot-path source files, unmodified, by path.

Roots, in order: $MMF_REFERENCE_ROOT, /root/reference (the build container), oracle/_ref (the files staged by
oracle/build_ref.py, which travel to the GPU box).  Never imported by the product or by the `-m gpu` tests.  Consumers:
oracle/make_golden.py (fixtures under tests/golden/) and the CPU arms of bench.py (`--impl reference`, `cpu_baseline`),
which time the reference's own implementation on the host cores.

Why a loader: `import mmf` fails here (omegaconf / pytorch_lightning / iopath ... are not installed
and transformers is 5.5, outside the reference's <=4.10.1 pin - SURVEY.md 8c), but the hot-path
files only need (i) `transformers.modeling_bert` aliased to `transformers.models.bert.modeling_bert`
and (ii) permissive stubs for the `mmf.*` / omegaconf / lightning imports they never call on this
path.  The arithmetic executed is the reference's own, from its own files:
    mmf/modules/hf_layers.py, mmf/models/vilbert.py, mmf/modules/embeddings.py,
    mmf/models/mmbt.py, mmf/models/transformers/backends/huggingface.py
"""
import importlib.abc
import importlib.machinery
import importlib.util
import os
import sys
import types

def _default_root():
    env = os.environ.get("MMF_REFERENCE_ROOT")
    if env:
        return env
    if os.path.isdir("/root/reference/mmf"):
        return "/root/reference"
    return os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")


REF_ROOT = _default_root()

# NB: `transformers3` is deliberately NOT stubbed: the reference's `try: from transformers3 ...` must fail
# so that it falls back to the real (aliased) transformers.modeling_bert.
_STUB_PREFIXES = ("mmf", "omegaconf", "pytorch_lightning", "iopath", "termcolor", "torchtext", "lmdb")


class _Anything:
    """Permissive stand-in: attribute access, calls, decorators and subclassing all 'work'."""

    def __init__(self, *a, **k):
        pass

    def __call__(self, *a, **k):
        # decorator usage: @registry.register_x("name") -> returns identity decorator
        if len(a) == 1 and not k and (isinstance(a[0], type) or callable(a[0])) and not isinstance(a[0], str):
            return a[0]
        return _Anything()

    def __getattr__(self, name):
        if name.startswith("__") and name.endswith("__"):
            raise AttributeError(name)
        return _Anything()

    def __mro_entries__(self, bases):
        import torch
        return (torch.nn.Module,)

    def __iter__(self):
        return iter(())


class _StubModule(types.ModuleType):
    def __getattr__(self, name):
        if name.startswith("__") and name.endswith("__"):
            raise AttributeError(name)
        return _Anything()


class _StubFinder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    def find_spec(self, fullname, path=None, target=None):
        root = fullname.split(".")[0]
        if root in _STUB_PREFIXES and fullname not in _REAL:
            return importlib.machinery.ModuleSpec(fullname, self, is_package=True)
        return None

    def create_module(self, spec):
        m = _StubModule(spec.name)
        m.__path__ = []
        return m

    def exec_module(self, module):
        pass


_REAL = {}
_installed = False


def _install():
    global _installed
    if _installed:
        return
    if not os.path.isdir(os.path.join(REF_ROOT, "mmf")):
        raise RuntimeError("reference files not present under %s (run `python -m oracle.build_ref` where /root/reference "
                           "exists; oracle/_ref then travels with the snapshot)" % REF_ROOT)
    import transformers.models.bert.modeling_bert as mb
    sys.modules.setdefault("transformers.modeling_bert", mb)
    sys.meta_path.insert(0, _StubFinder())
    _installed = True


def load(relpath, modname):
    """Executes /root/reference/<relpath> as module `modname` under the stub hook."""
    _install()
    if modname in _REAL:
        return _REAL[modname]
    path = os.path.join(REF_ROOT, relpath)
    spec = importlib.util.spec_from_file_location(modname, path)
    mod = importlib.util.module_from_spec(spec)
    _REAL[modname] = mod
    sys.modules[modname] = mod
    spec.loader.exec_module(mod)
    return mod


def hf_layers():
    return load("mmf/modules/hf_layers.py", "mmf.modules.hf_layers")


def vilbert():
    return load("mmf/models/vilbert.py", "mmf.models.vilbert")


def embeddings():
    return load("mmf/modules/embeddings.py", "mmf.modules.embeddings")


def mmbt():
    hf_layers()
    return load("mmf/models/mmbt.py", "mmf.models.mmbt")


def hf_backend():
    hf_layers()
    return load("mmf/models/transformers/backends/huggingface.py", "mmf.models.transformers.backends.huggingface")


def encoders():
    """mmf/modules/encoders.py: needs the pre-4.0 module paths of transformers' Auto classes as aliases."""
    _install()
    import transformers.models.auto.configuration_auto as ca
    import transformers.models.auto.modeling_auto as ma
    sys.modules.setdefault("transformers.configuration_auto", ca)
    sys.modules.setdefault("transformers.modeling_auto", ma)
    hf_layers()
    embeddings()
    return load("mmf/modules/encoders.py", "mmf.modules.encoders")


def optimizers():
    """mmf/modules/optimizers.py: `adam_w` (transformers AdamW, or torch.optim.AdamW when transformers has none - the case
    here) and AdamWSkipParamsWithZeroGrad, whose step() carries the transformers arithmetic inside the reference tree."""
    return load("mmf/modules/optimizers.py", "mmf.modules.optimizers")


def uniter():
    hf_layers()
    return load("mmf/models/uniter.py", "mmf.models.uniter")


def lxmert():
    """mmf/models/lxmert.py; its BertSelfAttention(encoder_hidden_states=...) call needs the reference's JIT layer
    implementations installed (hf_layers().replace_with_jit(), as the reference's own constructors do)."""
    hf_layers()
    return load("mmf/models/lxmert.py", "mmf.models.lxmert")


def vit():
    """mmf/modules/vit.py: HF ViT blocks with BertSelfAttention inside (maskable) - ViTLayer / ViTEncoder / ViTModel"""
    hf_layers()
    return load("mmf/modules/vit.py", "mmf.modules.vit")


def vinvl():
    """mmf/models/vinvl.py: VinVLBase (text BertEmbeddings + projected region features -> HF BertEncoder)"""
    hf_layers()
    return load("mmf/models/vinvl.py", "mmf.models.vinvl")


def visual_bert():
    """mmf/models/visual_bert.py.  VisualBERTBase.__init__ ends in HF's init_weights(), which in transformers 5 needs the
    post_init() bookkeeping the <= 4.10-era class never did; callers that construct it neutralise that one call (weights
    are set explicitly by the golden script) - the forward under test is the reference's own."""
    hf_layers()
    embeddings()
    return load("mmf/models/visual_bert.py", "mmf.models.visual_bert")
