"""Device-resident replay ring: the slice of `ReplayMemory`'s interface the update path uses
(`push`, `push_batch`, `sample`, `__len__`, `reset`; pyrl/env/replay_buffer.py:183-231,297-322) with the transitions
stored in HBM, so `sample(B)` is one gather kernel straight into the update engine's batch buffer instead of a numpy
`take` per key, a host staging copy and a host-to-device transfer per update.

Sampling follows the reference's `OneStepTransition` with replacement (env/sampling_strategy.py:26-31,93-101): indices
come from a host `np.random.RandomState(seed).randint(0, len(self), size=B)`, so a seeded run draws the same
transitions as the reference ring holding the same data; only the index vector (8 B per sample) crosses PCIe.
Capacity is bounded by HBM: 38.6 KB per transition at the ManiSkill shapes, i.e. 3.9 GB for the default 100 k ring."""
import numpy as np
import torch

from ._lib import lib, stream_ptr
from .data import unwrap

_KEEP = ("obs", "next_obs", "actions", "rewards", "dones", "episode_dones")


def _flatten(d, prefix=""):
    out = {}
    for k, v in d.items():
        if isinstance(v, dict):
            out.update(_flatten(v, f"{prefix}{k}/"))
        else:
            out[f"{prefix}{k}"] = v
    return out


class DeviceBatch:
    """What `DeviceReplayMemory.sample` returns: the ring plus the sampled positions.  The agent hands it to the
    engine (`gather_into`); `to_host()` materialises the same batch as nested numpy arrays for inspection/tests."""

    def __init__(self, ring, index_host, index_dev):
        self.ring, self.index, self._index_dev = ring, index_host, index_dev

    def __len__(self):
        return len(self.index)

    def gather_into(self, engine):
        self.ring._gather(self._index_dev, engine)

    def to_host(self):
        out = {}
        for key, buf in self.ring._leaves.items():
            node, ks = out, key.split("/")
            for k in ks[:-1]:
                node = node.setdefault(k, {})
            node[ks[-1]] = buf[self._index_dev].cpu().numpy()
        return out


class DeviceReplayMemory:
    def __init__(self, capacity, device="cuda", seed=None, keys=_KEEP):
        self.capacity, self.device, self.keys = int(capacity), torch.device(device), tuple(keys)
        self.seed = np.random.randint(0, 2**32 - 1) if seed is None else seed
        self.np_random = np.random.RandomState(self.seed)
        self._leaves = None  # flattened key -> device tensor [capacity, ...]
        self._tables = {}    # engine id -> (src_ptrs, dst_ptrs, row_bytes) device tables
        self.reset()

    # ------------------------------------------------------------------ ReplayMemory surface
    def __len__(self):
        return min(self.running_count, self.capacity)

    def reset(self):
        self.position, self.running_count = 0, 0

    def push(self, item):
        """One transition (leaves without the leading batch dimension)."""
        add_dim = lambda d: {k: (add_dim(v) if isinstance(v, dict) else np.asarray(v)[None]) for k, v in d.items()}
        self.push_batch(add_dim(unwrap(item)))

    def push_batch(self, items):
        flat = {k: v for k, v in _flatten(unwrap(items)).items() if k.split("/")[0] in self.keys}
        n = len(next(iter(flat.values())))
        if n > self.capacity:
            flat, n = {k: v[: self.capacity] for k, v in flat.items()}, self.capacity
        if self._leaves is None:  # first push defines the per-transition shapes (replay_buffer.py:217-219)
            self._leaves = {}
            for k, v in flat.items():
                t = torch.as_tensor(np.asarray(v))
                dt = torch.uint8 if t.dtype == torch.bool else (torch.float32 if t.is_floating_point() else t.dtype)
                self._leaves[k] = torch.zeros((self.capacity,) + tuple(t.shape[1:]), dtype=dt, device=self.device)
        first = min(n, self.capacity - self.position)
        for k, v in flat.items():
            buf = self._leaves[k]
            t = torch.as_tensor(np.ascontiguousarray(v))
            t = t.to(torch.uint8) if t.dtype == torch.bool else t.to(buf.dtype)
            t = t.reshape((n,) + tuple(buf.shape[1:])).to(self.device, non_blocking=True)
            buf[self.position:self.position + first].copy_(t[:first])
            if first < n:  # wrap around (replay_buffer.py:220-225)
                buf[: n - first].copy_(t[first:])
        self.running_count += n
        self.position = (self.position + n) % self.capacity

    def sample(self, batch_size, **_):
        if len(self) == 0:
            return None
        index = self.np_random.randint(low=0, high=len(self), size=batch_size)  # OneStepTransition, with replacement
        index_dev = torch.from_numpy(index.astype(np.int64)).to(self.device, non_blocking=True)
        return DeviceBatch(self, index, index_dev)

    # ------------------------------------------------------------------ engine hand-over
    def _leaf_for(self, key):
        """engine leaf name -> ring leaf (`state` is stored as `agent` or `state` in the observation dict)."""
        if key in self._leaves:
            return self._leaves[key]
        if key.endswith("/state"):
            alt = key[: -len("state")] + "agent"
            if alt in self._leaves:
                return self._leaves[alt]
        if key in ("rewards", "dones", "actions"):
            return self._leaves[key]
        raise KeyError(f"replay ring has no leaf for {key}")

    def _gather(self, index_dev, engine):
        tab = self._tables.get(id(engine))
        if tab is None:
            src, dst, nbytes = [], [], []
            for key, shape, dt, _off, nb in engine._batch_layout:
                ring_leaf, eng_leaf = self._leaf_for(key), engine._device_leaf(key)
                row = nb // engine.B
                if ring_leaf.dtype != eng_leaf.dtype or ring_leaf[0].numel() * ring_leaf.element_size() != row:
                    raise ValueError(f"replay leaf {key}: {tuple(ring_leaf.shape[1:])} {ring_leaf.dtype} does not match the "
                                     f"engine's {tuple(eng_leaf.shape[1:])} {eng_leaf.dtype}")
                src.append(ring_leaf.data_ptr())
                dst.append(eng_leaf.data_ptr())
                nbytes.append(row)
            mk = lambda v: torch.tensor(v, dtype=torch.int64, device=self.device)
            tab = (mk(src), mk(dst), mk(nbytes), len(src))
            self._tables[id(engine)] = tab
        if len(index_dev) != engine.B:
            raise ValueError(f"sampled {len(index_dev)} transitions for an engine built for batches of {engine.B}")
        lib().gather_transitions(tab[0], tab[1], tab[2], tab[3], index_dev, engine.B, stream_ptr())
