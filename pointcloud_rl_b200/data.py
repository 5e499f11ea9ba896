"""Minimal batch containers with the slice of pyrl's GDict/DictArray interface this path touches
(pyrl/utils/data/dict_array.py): the agent only needs `memory.sample(B)` to return something whose
`.memory` (or the object itself) is a nested dict of arrays, and `to_torch`.  The reference's own
DictArray/GDict objects are accepted too (duck-typed on `.memory`)."""
import numpy as np
import torch


def unwrap(x):
    """nested dict of arrays from a GDict/DictArray-like object or a plain dict."""
    m = getattr(x, "memory", x)
    if isinstance(m, dict):
        return {k: unwrap(v) for k, v in m.items()}
    return m


class GDict:
    def __init__(self, item=None):
        self.memory = unwrap(item) if item is not None else {}

    def __getitem__(self, key):
        node = self.memory
        for k in key.split("/"):
            node = node[k]
        return node

    def __setitem__(self, key, value):
        node = self.memory
        ks = key.split("/")
        for k in ks[:-1]:
            node = node.setdefault(k, {})
        node[ks[-1]] = value

    def __contains__(self, key):
        try:
            self[key]
            return True
        except (KeyError, TypeError):
            return False

    def keys(self):
        return self.memory.keys()

    def _map(self, fn, node=None):
        node = self.memory if node is None else node
        return {k: self._map(fn, v) if isinstance(v, dict) else fn(v) for k, v in node.items()}

    def to_torch(self, device="cpu", non_blocking=False, wrapper=True):
        out = self._map(lambda v: torch.as_tensor(v).to(device, non_blocking=non_blocking))
        return type(self)(out) if wrapper else out

    def to_numpy(self, wrapper=True):
        out = self._map(lambda v: v.detach().cpu().numpy() if torch.is_tensor(v) else np.asarray(v))
        return type(self)(out) if wrapper else out


class DictArray(GDict):
    """GDict whose leaves share the leading (batch) dimension."""

    def __init__(self, item=None, capacity=None):
        super().__init__(item)
        sizes = set()
        self._map(lambda v: sizes.add(len(v)))
        if len(sizes) > 1:
            raise ValueError(f"leaves disagree on the leading dimension: {sorted(sizes)}")
        self.capacity = sizes.pop() if sizes else (capacity or 0)

    def __len__(self):
        return self.capacity

    def take(self, index):
        return DictArray(self._map(lambda v: v[index]))


class ArrayMemory:
    """A host-numpy replay stand-in: `sample(n)` draws uniform-with-replacement indices like
    OneStepTransition (env/sampling_strategy.py:26-31,93-101) over a fixed DictArray."""

    def __init__(self, data, seed=0):
        self.data = data if isinstance(data, DictArray) else DictArray(data)
        self.rng = np.random.RandomState(seed)

    def __len__(self):
        return len(self.data)

    def sample(self, batch_size):
        return self.data.take(self.rng.randint(0, len(self.data), size=batch_size))


class FixedBatchMemory:
    """`sample(n)` always returns the same batch (parity tests / benchmarks)."""

    def __init__(self, batch):
        self.batch = batch

    def sample(self, batch_size):
        return DictArray(self.batch)
