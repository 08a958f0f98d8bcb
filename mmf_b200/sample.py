"""Minimal mirror of `mmf.common.sample.SampleList` (mmf/common/sample.py:69-397) for the keys this path consumes:
an OrderedDict of batched tensors with attribute access, nested dict fields (`image_info_0.max_features`),
`.to(device)`, `.pin_memory()`, `get_batch_size()`, `fields()`."""
from collections import OrderedDict

import torch


class SampleList(OrderedDict):
    def __getattr__(self, key):
        try:
            return self[key]
        except KeyError:
            raise AttributeError(key)

    def __setattr__(self, key, value):
        self[key] = value

    def fields(self):
        return list(self.keys())

    def get_batch_size(self):
        for v in self.values():
            if isinstance(v, torch.Tensor):
                return v.shape[0]
        return 0

    def _map(self, fn):
        out = SampleList()
        for k, v in self.items():
            if isinstance(v, torch.Tensor):
                out[k] = fn(v)
            elif isinstance(v, dict):
                out[k] = type(v)((kk, fn(vv) if isinstance(vv, torch.Tensor) else vv) for kk, vv in v.items())
            else:
                out[k] = v
        return out

    def to(self, device, non_blocking=True):
        return self._map(lambda t: t.to(device, non_blocking=non_blocking))

    def pin_memory(self):
        return self._map(lambda t: t.pin_memory())

    def to_dict(self):
        return dict(self)
