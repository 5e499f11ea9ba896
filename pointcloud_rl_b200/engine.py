"""Update engine: device-resident state + kernel sequencing for one SAC / DrQ `update_parameters`.

Mirrors, step for step, pyrl/methods/mfrl/sac.py:103-214 and drq.py:46-165 (ordered spec: SURVEY.md
Appendix B), but every arithmetic step is a libpcrl kernel launched through the C ABI
(include/pcrl.h).  PyTorch is used for device memory, streams and (dist.py) NCCL only -- there is no
autograd and no torch math on this path, and no CPU fallback.

Redundant work the reference does is removed without changing the maths: the 6 PointNet forwards per
update collapse to the 3 distinct ones (next_obs, obs with grad, first-aug obs with post-step weights),
and the two Q heads' feature gradients are summed before the single PointNet backward.
"""
import ctypes
import os
from dataclasses import dataclass, field
from typing import Dict, Optional, Tuple

import numpy as np
import torch

from ._lib import lib, stream_ptr

_NO_FORK = os.environ.get("PCRL_NO_FORK", "")
S_NAMES = [
    "critic_loss", "max_critic_abs_err", "q", "q_target", "critic_grad_sq", "actor_loss", "alpha_loss",
    "entropy", "actor_grad_sq", "alpha", "alpha_grad",
]
NUM_SCALARS = 16
AUG_KINDS = {None: 0, "none": 0, "jitter": 1, "rot": 2, "shift": 3, "downsample": 4, "colorjitter": 5}
MLP_KEYS = ["w0", "b0", "w1", "b1", "w2", "b2"]
PN_KEYS = ["pn.w0", "pn.b0", "pn.w1", "pn.g1", "pn.be1", "pn.w2", "pn.g2", "pn.be2", "pn.wf", "pn.bf", "pn.gf", "pn.bef"]


def _align(n, a):
    return (n + a - 1) // a * a


SCALAR_RING = 8  # asynchronous scalar results in flight (pinned slots)


class LazyScalars(dict):
    """The dict `update_parameters` returns (sac.py:140-203: name -> python float), filled in on first access.  The
    values were copied device->host behind the update; touching the dict (indexing, iteration, len, repr, ...) waits for
    that copy.  A caller that reads it right away sees exactly the synchronous behaviour; a training loop that logs the
    previous update's scalars after launching the next one never blocks on the GPU."""

    __slots__ = ("_engine", "_slot", "_event", "_updates")

    def __init__(self, engine, slot, event, updates):
        super().__init__()
        self._engine, self._slot, self._event, self._updates = engine, slot, event, updates

    def _resolve(self):
        eng = self._engine
        if eng is None:
            return
        pending = eng._scalar_pending
        while pending and pending[0] is not self:  # earlier updates first (alpha is carried from one to the next)
            pending[0]._resolve()
        self._event.synchronize()
        s = self._slot.numpy().astype(np.float64)
        self._engine = None
        if pending and pending[0] is self:
            pending.pop(0)
        dict.update(self, eng._scalars_dict(s, self._updates))

    @property
    def ready(self):
        """True once the values are on the host (never blocks)."""
        return self._engine is None or (self._engine._scalar_pending[0] is self and self._event.query())


def _resolving(name):
    base = getattr(dict, name)

    def method(self, *a, **k):
        self._resolve()
        return base(self, *a, **k)

    method.__name__ = name
    return method


for _n in ("__getitem__", "__iter__", "__len__", "__contains__", "__repr__", "__eq__", "__ne__", "__reversed__", "__or__",
           "__ror__", "__setitem__", "__delitem__", "get", "items", "keys", "values", "copy", "pop", "popitem", "setdefault",
           "update", "__str__", "__bool__"):
    if hasattr(dict, _n):
        setattr(LazyScalars, _n, _resolving(_n))
LazyScalars.__bool__ = lambda self: len(self) > 0
LazyScalars.__reduce__ = lambda self: (dict, (dict(self.items()),))


@dataclass
class PathSpec:
    """Shapes of the path (B, N, C, c=(c1,c2,c3), D, S, A)."""

    n_points: int
    action_dim: int
    state_dim: int = 0
    has_rgb: bool = True
    rgb_u8: bool = True
    n_pos: int = 0
    n_seg: int = 0
    widths: Tuple[int, int, int] = (128, 128, 256)
    out_dim: int = 128
    hidden: Tuple[int, int] = (1024, 1024)
    ln_eps: float = 1e-6
    head_ln_eps: float = 1e-5

    @property
    def C(self):
        return 3 + (3 if self.has_rgb else 0) + self.n_pos + self.n_seg

    @property
    def CP(self):
        return 8 if self.C <= 8 else 16

    @property
    def NP(self):
        return _align(self.n_points, 128)

    def param_shapes(self):
        c1, c2, c3 = self.widths
        D, S, A = self.out_dim, self.state_dim, self.action_dim
        h1, h2 = self.hidden
        pn = {
            "pn.w0": (c1, self.C), "pn.b0": (c1,), "pn.w1": (c2, c1), "pn.g1": (c2,), "pn.be1": (c2,),
            "pn.w2": (c3, c2), "pn.g2": (c3,), "pn.be2": (c3,), "pn.wf": (D, c3), "pn.bf": (D,), "pn.gf": (D,),
            "pn.bef": (D,),
        }

        def mlp(net, din, dout):
            return {
                f"{net}.w0": (h1, din), f"{net}.b0": (h1,), f"{net}.w1": (h2, h1), f"{net}.b1": (h2,),
                f"{net}.w2": (dout, h2), f"{net}.b2": (dout,),
            }

        groups = {
            "critic": {**pn, **mlp("q0", D + S + A, 1), **mlp("q1", D + S + A, 1)},
            "actor": mlp("actor", D + S, 2 * A),
            "alpha": {"log_alpha": (1,)},
            "target": {**mlp("tq0", D + S + A, 1), **mlp("tq1", D + S + A, 1)},
        }
        return groups


@dataclass
class HyperParams:
    algo: str = "sac"
    gamma: float = 0.99
    reward_scale: float = 1.0
    num_aug: int = 1
    aug: Optional[str] = None
    aug_lo: float = 0.0
    aug_hi: float = 0.0
    aug_axes: int = 7  # shift only: bit i set = axis i is translated (include/pcrl.h, PCRL_AUG_SHIFT_AXES)
    aug_color: Tuple[float, float, float, float] = (0.0, 0.0, 0.0, 0.0)  # colorjitter: brightness, contrast, saturation, hue
    tau: float = 0.01
    actor_update_interval: int = 2
    target_update_interval: int = 2
    lr: float = 1e-3
    actor_lr: float = 1e-3
    alpha_lr: float = 1e-3
    betas: Tuple[float, float] = (0.9, 0.999)
    actor_betas: Tuple[float, float] = (0.9, 0.999)
    alpha_betas: Tuple[float, float] = (0.5, 0.999)
    adam_eps: float = 1e-8
    log_std_bound: Tuple[float, float] = (-10.0, 2.0)
    head_scale: float = 1.0
    head_bias: float = 0.0
    target_entropy: Optional[float] = None
    ignore_dones: bool = False
    automatic_alpha_tuning: bool = True


class ParamLayout:
    """One flat fp32 buffer [critic | actor | alpha | target]; every tensor 16-byte aligned so the fused
    Adam / Polyak kernels run float4-wide over whole groups."""

    def __init__(self, spec: PathSpec):
        self.entries: Dict[str, Tuple[int, Tuple[int, ...]]] = {}
        self.group_range: Dict[str, Tuple[int, int]] = {}
        off = 0
        for gname, tensors in spec.param_shapes().items():
            g0 = off
            for name, shape in tensors.items():
                self.entries[name] = (off, shape)
                off += _align(int(np.prod(shape)), 4)
            self.group_range[gname] = (g0, off)
        self.total = off
        self.trainable = self.group_range["alpha"][1]  # critic | actor | alpha carry grads + Adam state
        c0 = self.group_range["critic"][0]
        self.q_range = (self.entries["q0.w0"][0] - c0, self.group_range["critic"][1] - c0)  # inside critic group
        assert self.q_range[1] - self.q_range[0] == self.group_range["target"][1] - self.group_range["target"][0]

    def views(self, flat: torch.Tensor, names=None):
        out = {}
        for name, (off, shape) in self.entries.items():
            if off + int(np.prod(shape)) > flat.numel():
                continue
            if names is None or name in names:
                out[name] = flat[off: off + int(np.prod(shape))].view(*shape)
        return out


class UpdateEngine:
    def __init__(self, spec: PathSpec, hp: HyperParams, batch_size: int, device="cuda:0", precision="fp32",
                 seed: int = 0, fwd_chunk_clouds: int = 64):
        assert precision in ("fp32", "tf32", "bf16")
        self.L = lib()
        self.spec, self.hp, self.B = spec, hp, int(batch_size)
        self.k = hp.num_aug if hp.algo == "drq" else 1
        self.R = self.B * self.k
        self.device = torch.device(device)
        self.precision = precision
        # "bf16" = the fast mode: fused bf16 tcgen05 PointNet forward + TF32 tcgen05 GEMMs for the MLP heads and the
        # compacted backward; "tf32" = the reference-precision tier on tensor cores: every GEMM of the path (PointNet
        # layers 1-2 included) on the TF32 tcgen05 kernel, all statistics / max / argmax exact fp32; "fp32" = the
        # parity mode: every product on the exact-fp32 FFMA kernels
        self.tf32 = 1 if precision in ("bf16", "tf32") else 0
        if precision == "tf32":
            fwd_chunk_clouds = min(fwd_chunk_clouds, 32)  # a chunk's activations (h0|h1|h2) stay inside the 126 MB L2
        self.seed = int(seed)
        self.layout = ParamLayout(spec)
        dev = self.device
        if dev.type == "cuda":
            # the library's per-device context (SM count, the backward's internal side stream + events): created here,
            # explicitly, rather than on first use inside a kernel call
            self.ctx = int(self.L.create(dev.index if dev.index is not None else torch.cuda.current_device()))
            if not self.ctx:
                raise RuntimeError(f"pcrl_create failed: {self.L.last_error()}")
        f32 = dict(dtype=torch.float32, device=dev)
        self.params = torch.zeros(self.layout.total, **f32)
        self.grads = torch.zeros(self.layout.trainable, **f32)
        self.adam_m = torch.zeros(self.layout.trainable, **f32)
        self.adam_v = torch.zeros(self.layout.trainable, **f32)
        self.p = self.layout.views(self.params)
        self.g = self.layout.views(self.grads)
        self.steps = torch.zeros(4, dtype=torch.int32, device=dev)  # critic, actor, alpha Adam step counters
        self.alpha_dev = torch.zeros(1, **f32)
        self.scalars = torch.zeros(NUM_SCALARS, **f32)
        self.scalars_host = torch.zeros(NUM_SCALARS, dtype=torch.float32).pin_memory() if dev.type == "cuda" else None
        self.counter = torch.zeros(1, dtype=torch.int64, device=dev)  # Philox offset, bumped every update
        self.world_size = 1
        self.allreduce = None  # set by dist.attach(): callable(flat_grad_view)
        self._alloc_workspace(fwd_chunk_clouds)

    def rebind_grads(self, new_grads):
        """Move the flat gradient buffer (dist.PeerAllReduce re-allocates it as symmetric memory).  Only before the first
        graph capture: captured kernels hold the old addresses."""
        if self._graphs:
            raise RuntimeError("rebind_grads after graph capture")
        assert new_grads.numel() == self.grads.numel() and new_grads.dtype == self.grads.dtype
        new_grads.copy_(self.grads)
        self.grads = new_grads
        self.g = self.layout.views(self.grads)
        self.w["dlog_alpha"] = self.g["log_alpha"]

    # ------------------------------------------------------------------ buffers
    def _alloc_workspace(self, fwd_chunk_clouds):
        sp, B, R, dev = self.spec, self.B, self.R, self.device
        f32 = dict(dtype=torch.float32, device=dev)
        N, NP, CP, A, S, D = sp.n_points, sp.NP, sp.CP, sp.action_dim, sp.state_dim, sp.out_dim
        c1, c2, c3 = sp.widths
        h1, h2 = sp.hidden
        u8 = dict(dtype=torch.uint8, device=dev)
        Kp = (D + S + 3) // 4 * 4
        self.actor_w0p = torch.zeros(h1, Kp, **f32) if (self.tf32 and Kp != D + S and Kp <= D + S + A) else None
        # The replay sample lives in ONE contiguous device buffer (leaves are 256-byte aligned views), so a batch
        # moves host -> landing slot -> here with a single copy each instead of one per leaf.
        leaves = []
        for which in ("obs", "next_obs"):
            leaves.append((f"{which}/xyz", (B, 3, N), torch.float32))
            if sp.has_rgb:
                leaves.append((f"{which}/rgb", (B, 3, N), torch.uint8 if sp.rgb_u8 else torch.float32))
            if sp.n_pos:
                leaves.append((f"{which}/pos_encoding", (B, sp.n_pos, N), torch.uint8))
            if sp.n_seg:
                leaves.append((f"{which}/seg", (B, sp.n_seg, N), torch.uint8))
            if S:
                leaves.append((f"{which}/state", (B, S), torch.float32))
        leaves += [("actions", (B, A), torch.float32), ("rewards", (B,), torch.float32), ("dones", (B,), torch.uint8)]
        self._batch_layout, off = [], 0
        for key, shape, dt in leaves:
            nbytes = int(np.prod(shape)) * torch.empty((), dtype=dt).element_size()
            self._batch_layout.append((key, shape, dt, off, nbytes))
            off += (nbytes + 255) // 256 * 256
        self._batch_bytes = off
        self.raw_flat = torch.zeros(off, **u8)
        self.raw = {"obs": {}, "next_obs": {}}
        for key, view in self._batch_views(self.raw_flat).items():
            if "/" in key:
                a, b = key.split("/")
                self.raw[a][b] = view
            else:
                self.raw[key] = view
        self._pinned = None

        w = {}
        bf16 = self.precision == "bf16"
        if self.hp.algo == "drq" and self.hp.aug == "colorjitter":
            if not (sp.has_rgb and sp.rgb_u8):
                raise NotImplementedError("ColorJitterPoints needs uint8 point colours (torchvision's uint8 path)")
            w["rgb_aug_next"] = torch.zeros(B, 3, N, **u8)
            w["rgb_aug_obs"] = torch.zeros(B, 3, N, **u8)
        if self.hp.algo == "drq" and self.hp.aug == "downsample":
            w["ds_map_next"] = torch.zeros(N, dtype=torch.int32, device=dev)
            w["ds_map_obs"] = torch.zeros(N, dtype=torch.int32, device=dev)
        for name, rows in (("next", R), ("obs", R), ("pi", B)):
            if not bf16 or name == "obs":  # tensor-core path: only the critic's backward reads the fp32 staging
                w[f"xf_{name}"] = torch.zeros(rows, NP, CP, **f32)
            if bf16:
                w[f"xh_{name}"] = torch.zeros(rows * NP * 16, dtype=torch.bfloat16, device=dev)
            w[f"pooled_{name}"] = torch.zeros(rows, c3, **f32)
            w[f"cat_{name}"] = torch.zeros(rows, D + S + A, **f32)
            w[f"z_{name}"] = torch.zeros(rows, D, **f32)
            w[f"nlp_{name}"] = torch.zeros(rows, **f32)
            w[f"eps_{name}"] = torch.zeros(rows, A, **f32)
            w[f"out_{name}"] = torch.zeros(rows, 2 * A, **f32)
            w[f"q_{name}"] = torch.zeros(rows, 2, **f32)
        w["argmax_obs"] = torch.zeros(R, c3, dtype=torch.int32, device=dev)
        w["xhat_obs"] = torch.zeros(R, D, **f32)
        w["rstd_obs"] = torch.zeros(R, **f32)
        w["y"] = torch.zeros(R, **f32)
        w["dq"] = torch.zeros(R, 2, **f32)
        w["dz"] = torch.zeros(R, D, **f32)
        w["dpooled"] = torch.zeros(R, c3, **f32)
        w["dout"] = torch.zeros(B, 2 * A, **f32)
        w["dlog_alpha"] = self.g["log_alpha"]
        # MLP activations: [net][layer]; no-grad passes reuse "tmp"
        for net in ("tmp", "tmp2", "q0", "q1", "actor"):
            w[f"h1_{net}"] = torch.zeros(R, h1, **f32)
            w[f"h2_{net}"] = torch.zeros(R, h2, **f32)
        for sc in ("a", "b"):  # backward scratch of the two Q heads, which run on parallel streams
            w[f"dh1_{sc}"] = torch.zeros(R, h1, **f32)
            w[f"dh2_{sc}"] = torch.zeros(R, h2, **f32)
        w["dx0"] = torch.zeros(R, D + S + A, **f32)
        w["dx1"] = torch.zeros(R, D + S + A, **f32)
        chunk = max(1, min(R, fwd_chunk_clouds))
        ws_query = self.L.pointnet_fwd_tf32_workspace if self.precision == "tf32" else self.L.pointnet_fwd_f32_workspace
        self.fwd_ws_bytes = int(ws_query(chunk, NP, c1, c2, c3))
        self.bwd_ws_bytes = int(self.L.pointnet_bwd_workspace(R, NP, c1, c2, c3, CP))
        w["scratch"] = torch.zeros(max(self.fwd_ws_bytes, self.bwd_ws_bytes), dtype=torch.uint8, device=dev)
        if not bf16:
            # the target branch's encode(next) runs on a forked stream beside encode(obs): it needs its own forward
            # workspace (the bf16 path keeps its intermediates on chip and has per-branch pool_keys_* buffers)
            w["scratch_next"] = torch.zeros(self.fwd_ws_bytes, dtype=torch.uint8, device=dev)
        if bf16:
            w["wpack"] = torch.zeros(int(self.L.pointnet_wpack_bytes(c1, c2, c3)), dtype=torch.uint8, device=dev)
            for name in ("next", "obs", "pi"):
                w[f"pool_keys_{name}"] = torch.zeros(R * c3, dtype=torch.int64, device=dev)
        self.w = w
        self._graphs = {}
        self._noise_static = {}
        self.graph_calls = {}
        # side streams 0-2 carry branches of the critical chain (target branch, second Q head): high priority, like the
        # capture stream; 3-8 carry the weight-gradient GEMMs that only feed the optimizer: default (low) priority, so
        # the block scheduler hands free SMs to the dX chain first
        # 9 carries the packing of the backward's weight image (joined right before the PointNet backward)
        self._side = ([torch.cuda.Stream(device=dev, priority=-1 if i < 3 else 0) for i in range(10)]
                      if dev.type == "cuda" else [])
        self._capture_stream = torch.cuda.Stream(device=dev, priority=-1) if dev.type == "cuda" else None
        self._landing = None

    # ------------------------------------------------------------------ parameters
    def load_params(self, params: Dict[str, torch.Tensor]):
        """Copies oracle/reference-named tensors (pn.*, actor.*, q0.*, q1.*, [tq0.*, tq1.*], log_alpha)."""
        for name, view in self.p.items():
            src = params.get(name)
            if src is None and name.startswith("tq"):
                src = params[name[1:]]  # hard_update(target, critic), builder.py:43
            if src is None:
                raise KeyError(f"missing parameter {name}")
            view.copy_(torch.as_tensor(src, dtype=torch.float32).reshape(view.shape))
        self.refresh_alpha()

    def export_params(self):
        return {k: v.detach().clone().cpu() for k, v in self.p.items()}

    def refresh_alpha(self):
        with torch.cuda.device(self.device):
            self.L.refresh_alpha(self.p["log_alpha"], self.alpha_dev, self.scalars, stream_ptr())

    # ------------------------------------------------------------------ batch upload (host -> device)
    def upload_batch(self, batch):
        """batch: dict(obs, next_obs, actions, rewards, dones) of numpy arrays / torch CPU tensors (the
        reference's `memory.sample(B)` layout, replay_buffer.py:297-322).  Every leaf is copied into its slot of ONE
        pinned staging buffer (multi-threaded memcpy inside the library) and its host->device copy is enqueued at once
        on a COPY stream into a landing buffer in HBM, so the DMA of leaf i overlaps the host memcpy of leaf i+1 and the
        whole transfer overlaps the previous update still running on the compute stream; the compute stream then waits
        for the DMA and adopts the landing buffer with one device-to-device copy (10 MB, ~7 us).  Pinned and landing
        buffers alternate (two each) so a call never overwrites bytes an earlier copy may still be reading."""
        if self._pinned is None:
            self._pinned = [torch.empty(self._batch_bytes, dtype=torch.uint8).pin_memory() for _ in range(2)]
            self._pinned_np = [{k: v.numpy() for k, v in self._batch_views(p).items()} for p in self._pinned]
            self._pinned_ev = [None, None]
            self._pinned_i = 0
            self._landing = [torch.empty_like(self.raw_flat) for _ in range(2)]
            self._landing_free = [None, None]
            self._copy_stream = torch.cuda.Stream(device=self.device)
            n = len(self._batch_layout)
            self._leaf_srcs = (ctypes.c_void_p * n)()
            self._leaf_offs = (ctypes.c_int64 * n)(*[off for *_, off, _nb in self._batch_layout])
            self._leaf_sizes = (ctypes.c_int64 * n)(*[nb for *_, nb in self._batch_layout])
            # staging threads: up to 4, but never more than this rank's share of the host cores (8 ranks on a 16-core
            # box get one each: oversubscribed memcpy threads slow every rank's host path down)
            ranks_here = int(os.environ.get("LOCAL_WORLD_SIZE", os.environ.get("WORLD_SIZE", "1")) or 1)
            self._copy_threads = max(1, min(4, (os.cpu_count() or 4) // max(1, ranks_here) - 1))
        i = self._pinned_i = self._pinned_i ^ 1
        if self._pinned_ev[i] is not None:
            self._pinned_ev[i].synchronize()
        host, views, land = self._pinned[i], self._pinned_np[i], self._landing[i]
        base = host.data_ptr()
        with torch.cuda.device(self.device):
            cs, main = self._copy_stream, torch.cuda.current_stream()
            if self._landing_free[i] is not None:
                cs.wait_event(self._landing_free[i])  # the adopt copy that last read this landing buffer
            keep = []  # source arrays stay referenced until the call returns
            for j, (key, _shape, _dt, off, nbytes) in enumerate(self._batch_layout):
                src = self._host_leaf(batch, key)
                dst = views[key]
                if src.flags.c_contiguous and src.dtype.itemsize == dst.dtype.itemsize and src.nbytes == nbytes and (
                        src.dtype == dst.dtype or src.dtype == np.bool_):
                    # same bytes (float32 -> float32, uint8 / bool -> uint8): the library's multi-threaded memcpy
                    keep.append(src)
                    self._leaf_srcs[j] = src.ctypes.data
                else:
                    np.copyto(dst, src.reshape(dst.shape), casting="unsafe")  # f64 -> f32, strided sources, ...
                    self._leaf_srcs[j] = None
            # one foreign call: stage every leaf into the pinned buffer and enqueue its DMA on the copy stream at once
            self.L.upload_leaves(base, land, ctypes.addressof(self._leaf_srcs), ctypes.addressof(self._leaf_offs),
                                 ctypes.addressof(self._leaf_sizes), len(self._batch_layout), self._copy_threads,
                                 cs.cuda_stream)
            del keep
            ev = torch.cuda.Event()
            ev.record(cs)
            self._pinned_ev[i] = ev
            main.wait_event(ev)
            self.raw_flat.copy_(land, non_blocking=True)
            free = torch.cuda.Event()
            free.record(main)
            self._landing_free[i] = free
        return sum(n for *_, n in self._batch_layout)

    @staticmethod
    def _host_leaf(batch, key):
        if "/" in key:
            which, leaf = key.split("/")
            obs = batch[which]
            v = obs[leaf] if leaf in obs else obs["agent" if leaf == "state" else leaf]
        else:
            v = batch[key]
        if torch.is_tensor(v):
            v = v.detach().cpu().numpy()
        return np.asarray(v)

    def _batch_views(self, flat):
        """Typed leaf views (flattened keys) into a byte buffer laid out like `raw_flat`."""
        out = {}
        for key, shape, dt, off, nbytes in self._batch_layout:
            out[key] = flat[off:off + nbytes].view(dt).view(shape)
        return out

    def _device_leaf(self, key):
        if "/" in key:
            a, b = key.split("/")
            return self.raw[a][b]
        return self.raw[key]

    def _flatten_batch(self, batch):
        out = {}
        for which in ("obs", "next_obs"):
            obs = batch[which]
            for k, v in obs.items():
                kk = "state" if k in ("state", "agent") else k
                if kk not in self.raw[which]:
                    continue
                t = torch.as_tensor(np.ascontiguousarray(v) if isinstance(v, np.ndarray) else v)
                tgt = self.raw[which][kk]
                if t.dtype == torch.bool:
                    t = t.to(torch.uint8)
                out[f"{which}/{kk}"] = t.to(tgt.dtype).reshape(tgt.shape)
        out["actions"] = torch.as_tensor(batch["actions"]).float().reshape(self.B, -1)
        out["rewards"] = torch.as_tensor(batch["rewards"]).float().reshape(self.B)
        out["dones"] = torch.as_tensor(np.asarray(batch["dones"]).astype(np.uint8)).reshape(self.B)
        return out

    # ------------------------------------------------------------------ building blocks
    def _stage(self, which, name, repeat, aug_kind, noise, stream_id, st):
        sp, raw = self.spec, self.raw[which]
        rgb = raw.get("rgb")
        if aug_kind == 5:
            # ColorJitterPoints: one parameter set per call is shared by the whole batch, so the num_aug copies of a
            # sample are identical: jitter the B source clouds once, the staging kernel repeats them
            c = self.hp.aug_color
            rgb = self.w[f"rgb_aug_{name}"]
            self.L.color_jitter_points(raw["rgb"], self.B, sp.n_points, noise, float(c[0]), float(c[1]), float(c[2]),
                                       float(c[3]), self.seed, self.counter, stream_id, rgb, st)
            aug_kind, noise = 0, None
        self.L.stage_points(
            raw["xyz"], rgb, int(sp.rgb_u8), raw.get("pos_encoding"), sp.n_pos, raw.get("seg"), sp.n_seg,
            self.B, sp.n_points, repeat, aug_kind, float(self.hp.aug_lo), float(self.hp.aug_hi), noise, self.seed,
            self.counter, stream_id, self.w.get(f"xf_{name}"), self.w.get(f"xh_{name}"), sp.CP, st)

    def _encode(self, name, rows, want_argmax, st):
        """PointNet per-point MLP + max-pool (pointnet.py:147-151) then final_mlp (pointnet.py:110,152-153):
        writes the feature into cat_<name>[:, :D]."""
        sp, w, p = self.spec, self.w, self.p
        c1, c2, c3 = sp.widths
        self._encode_points(name, rows, want_argmax, st)
        D = sp.out_dim
        cat = w[f"cat_{name}"]
        self.L.linear_fwd(w[f"pooled_{name}"], c3, p["pn.wf"], p["pn.bf"], w[f"z_{name}"], D, rows, c3, D, 0, self.tf32, st)
        save = want_argmax
        self.L.layernorm_fwd(w[f"z_{name}"], p["pn.gf"], p["pn.bef"], cat, cat.stride(0),
                             w["xhat_obs"] if save else None, w["rstd_obs"] if save else None, rows, D,
                             sp.head_ln_eps, st)

    def _encode_points(self, name, rows, want_argmax, st):
        """The fused per-point MLP + LayerNorm + ReLU + max-pool (+argmax) kernel: staged points -> pooled_<name>."""
        sp, w, p = self.spec, self.w, self.p
        c1, c2, c3 = sp.widths
        argmax = w["argmax_obs"] if want_argmax else None
        if self.precision == "bf16":
            if name == "pi" and self.k > 1:  # first augmentation of every sample, read in place from the obs staging
                self.L.pointnet_fwd_bf16_strided(w["xh_obs"], rows, self.k, sp.n_points, sp.NP, w["wpack"], c1, c2, c3,
                                                 sp.ln_eps, w["pool_keys_pi"], w["pooled_pi"], argmax, st)
            else:
                self.L.pointnet_fwd_bf16(w[f"xh_{name}"], rows, sp.n_points, sp.NP, w["wpack"], c1, c2, c3, sp.ln_eps,
                                         w[f"pool_keys_{name}"], w[f"pooled_{name}"], argmax, st)
        else:
            fwd = self.L.pointnet_fwd_tf32 if self.precision == "tf32" else self.L.pointnet_fwd_f32
            fwd(w[f"xf_{name}"], rows, sp.n_points, sp.NP, sp.CP, sp.C, p["pn.w0"], p["pn.b0"], p["pn.w1"], p["pn.g1"],
                p["pn.be1"], p["pn.w2"], p["pn.g2"], p["pn.be2"], c1, c2, c3, sp.ln_eps, w[f"pooled_{name}"], argmax,
                w["scratch_next" if name == "next" else "scratch"], self.fwd_ws_bytes, st)

    def dominant_kernel_name(self):
        return {"bf16": "pointnet_fwd_tc2_kernel (fused tcgen05 forward)",
                "tf32": "pointnet_fwd_tf32 chain (TF32 tcgen05 GEMMs + LayerNorm + max-pool)",
                "fp32": "pointnet_fwd_f32 chain (exact FFMA)"}[self.precision]

    def _pack_weights(self, st, which=3):
        """bf16 path: re-pack the (just updated) PointNet weights into the MMA-ready images.  which: 1 = the image the
        fused forward reads (critical path of every encode), 2 = the image the backward's recompute reads, 3 = both."""
        if self.precision != "bf16":
            return
        sp, p = self.spec, self.p
        c1, c2, c3 = sp.widths
        self.L.pointnet_pack_weights_part(p["pn.w0"], p["pn.b0"], p["pn.w1"], p["pn.g1"], p["pn.be1"], p["pn.w2"],
                                          p["pn.g2"], p["pn.be2"], sp.C, c1, c2, c3, int(sp.has_rgb and sp.rgb_u8), which,
                                          self.w["wpack"], st)

    def _mlp_fwd(self, net, x, K, M, out, ldo, nout, keep, st):
        p, (h1n, h2n) = self.p, self.spec.hidden
        h1, h2 = self.w[f"h1_{keep}"], self.w[f"h2_{keep}"]
        ldx = x.stride(0)
        w0 = p[f"{net}.w0"]
        if net == "actor" and self.actor_w0p is not None:
            # TMA needs a 16-byte row pitch: the actor's first layer reads a zero-padded copy of its [h, D+S] weight
            # (the extra input columns are the first action columns of `cat`, multiplied by exact zeros)
            w0, K = self.actor_w0p, self.actor_w0p.shape[1]
        self.L.linear_fwd(x, ldx, w0, p[f"{net}.b0"], h1, h1n, M, K, h1n, 1, self.tf32, st)
        self.L.linear_fwd(h1, h1n, p[f"{net}.w1"], p[f"{net}.b1"], h2, h2n, M, h1n, h2n, 1, self.tf32, st)
        self.L.linear_fwd(h2, h2n, p[f"{net}.w2"], p[f"{net}.b2"], out, ldo, M, h2n, nout, 0, self.tf32, st)

    def _mlp_bwd(self, net, x, K, M, dout, lddo, nout, keep, dx, want_w, st, scratch="a", wstream=None, defer=False,
                 join=True):
        """Backward of one 3-layer head.  The data-gradient chain (dX of layer 3 -> 2 -> 1) is the critical path; the
        weight/bias gradients only feed the optimizer, so with `wstream` (three side-stream indices) they run on forked
        streams.  defer=False: each is forked as soon as its dY exists and all are joined before returning.
        defer=True: they are enqueued after the whole dX chain (they no longer take SMs away from it) and the side
        streams are returned; the caller joins them (`_join`) before the optimizer step."""
        p, g, (h1n, h2n) = self.p, self.g, self.spec.hidden
        h1, h2 = self.w[f"h1_{keep}"], self.w[f"h2_{keep}"]
        w = {"dh1": self.w[f"dh1_{scratch}"], "dh2": self.w[f"dh2_{scratch}"]}
        L = self.L
        sides = [self._side[i] for i in wstream] if (want_w and wstream is not None) else None
        main = torch.cuda.current_stream()
        pending = []

        def wgrad(xin, ldx, name, dy, lddy, Kin, Nout_):
            """dW/db of one layer (dy is complete on the main stream when this runs); every layer has its own side
            stream, so the three weight-gradient GEMMs of a head also overlap each other."""
            if not want_w:
                return
            side = sides[name] if sides is not None else None
            if side is None:
                L.linear_bwd(xin, ldx, p[f"{net}.w{name}"], dy, lddy, g[f"{net}.w{name}"], g[f"{net}.b{name}"], None, 0,
                             None, 0, M, Kin, Nout_, self.tf32, stream_ptr())
                return
            side.wait_stream(main)
            with torch.cuda.stream(side):
                L.linear_bwd(xin, ldx, p[f"{net}.w{name}"], dy, lddy, g[f"{net}.w{name}"], g[f"{net}.b{name}"], None, 0,
                             None, 0, M, Kin, Nout_, self.tf32, stream_ptr())

        def maybe(*a):
            if defer and sides is not None:
                pending.append(a)
            else:
                wgrad(*a)

        # each layer's dX GEMM applies the previous ReLU's backward in its epilogue (mask = saved post-activation)
        maybe(h2, h2n, 2, dout, lddo, h2n, nout)
        L.linear_bwd(h2, h2n, p[f"{net}.w2"], dout, lddo, None, None, w["dh2"], h2n, h2, h2n, M, h2n, nout, self.tf32,
                     stream_ptr())
        maybe(h1, h1n, 1, w["dh2"], h2n, h1n, h2n)
        L.linear_bwd(h1, h1n, p[f"{net}.w1"], w["dh2"], h2n, None, None, w["dh1"], h1n, h1, h1n, M, h1n, h2n, self.tf32,
                     stream_ptr())
        maybe(x, x.stride(0), 0, w["dh1"], h1n, K, h1n)
        if dx is not None:
            L.linear_bwd(x, x.stride(0), p[f"{net}.w0"], w["dh1"], h1n, None, None, dx, dx.stride(0), None, 0, M, K, h1n,
                         self.tf32, stream_ptr())
        for a in pending:
            wgrad(*a)
        if sides is None:
            return []
        if defer or not join:
            return sides  # [layer 0, layer 1, layer 2] weight-gradient streams: the caller joins them
        for side in sides:
            main.wait_stream(side)
        return []

    def _adam(self, group, idx, lr, betas, gradsq_slot, polyak, st):
        lo, hi = self.layout.group_range[group]
        hp = self.hp
        target, pb, pe = None, 0, 0
        if polyak:
            t0, t1 = self.layout.group_range["target"]
            target = self.params[t0:t1]
            pb, pe = self.layout.q_range
        self.L.adam_step(self.params[lo:hi], self.grads[lo:hi], self.adam_m[lo:hi], self.adam_v[lo:hi], hi - lo,
                         float(lr), float(betas[0]), float(betas[1]), float(hp.adam_eps), 1.0 / self.world_size,
                         self.steps[idx:], self.scalars[gradsq_slot:] if gradsq_slot is not None else None, target,
                         pb, pe, float(hp.tau), st)

    def _adam_part(self, group, rng, idx, lr, betas, gradsq_slot, first, polyak, st):
        """Adam over the sub-range `rng` (offsets inside the group) of an optimizer group; `first` = the call that bumps
        the step count and clears the gradient-norm accumulator.  polyak: the range IS the Q heads (target layout)."""
        g0 = self.layout.group_range[group][0]
        lo, hi = g0 + rng[0], g0 + rng[1]
        hp = self.hp
        target = None
        if polyak:
            t0, t1 = self.layout.group_range["target"]
            target = self.params[t0:t1]
        self.L.adam_step_part(self.params[lo:hi], self.grads[lo:hi], self.adam_m[lo:hi], self.adam_v[lo:hi], hi - lo,
                              float(lr), float(betas[0]), float(betas[1]), float(hp.adam_eps), 1.0 / self.world_size,
                              self.steps[idx:], int(first), self.scalars[gradsq_slot:], int(first), target, 0,
                              hi - lo if polyak else 0, float(hp.tau), st)

    # ------------------------------------------------------------------ the update
    # ------------------------------------------------------------------ stream forks
    def _fork(self, idx):
        """Side stream `idx`, ordered after everything enqueued so far on the current stream.  Independent chains of
        the update (target branch vs critic forward, the two Q heads) run on forked streams: most of their kernels
        fill only part of the 148 SMs, and inside a CUDA graph the forks become parallel branches."""
        if _NO_FORK == "1" or (_NO_FORK and str(idx) in _NO_FORK.split(",")):  # debugging aid: PCRL_NO_FORK=1 | "0,1"
            return torch.cuda.current_stream()
        if idx in (0, 1) and self.R * self.spec.NP >= (1 << 22):
            # >= 4 M points per encode (BASELINE config 5: 8.4 M): each encode fills all 148 SMs for milliseconds, so
            # running the target branch beside the critic forward buys nothing -- and the forked form of this graph
            # fails at replay with "unspecified launch failure" at 8.4 M points (B=512 x N=16384; B=256 runs; every
            # other fork may stay; clean under compute-sanitizer, which serialises the branches).  Root cause not
            # found: the branch runs in line at these sizes.
            return torch.cuda.current_stream()
        side = self._side[idx]
        side.wait_stream(torch.cuda.current_stream())
        return side

    @staticmethod
    def _join(*sides):
        cur = torch.cuda.current_stream()
        for s_ in sides:
            cur.wait_stream(s_)

    def update(self, updates: int, noise: Optional[Dict[str, torch.Tensor]] = None):
        """Enqueues one full update on the current stream of the engine's device.  `noise` (parity mode) injects the
        reference's random draws: jitter_obs/jitter_next, angle_obs/angle_next or shift_obs/shift_next, eps_next, eps_pi
        (device tensors)."""
        with torch.cuda.device(self.device):  # kernels launch on the CURRENT device: pin it to the engine's
            self._update(updates, noise)

    def _update(self, updates, noise):
        sp, hp, w, p, L = self.spec, self.hp, self.w, self.p, self.L
        ST = stream_ptr  # evaluated at every call site: forked sections run on their own stream
        B, R, k = self.B, self.R, self.k
        D, S, A = sp.out_dim, sp.state_dim, sp.action_dim
        c1, c2, c3 = sp.widths
        noise = noise or {}
        aug = AUG_KINDS[hp.aug] if hp.algo == "drq" else 0
        nkey = {1: "jitter", 2: "angle", 3: "shift", 4: "keep", 5: "cj"}.get(aug)
        if aug == 3:
            aug |= (int(hp.aug_axes) & 7) << 8
        if aug == 4:
            # RandomDownSample: the stage kernel takes a source map [N].  Parity mode: built from the injected kept
            # indices; otherwise drawn on the device (aug_lo = drop_ratio, aug_hi = fixed_ratio)
            maps = {}
            for which, sid in (("next", 1), ("obs", 0)):
                m = w[f"ds_map_{which}"]
                keep = noise.get(f"keep_{which}")
                if keep is not None:
                    if not torch.cuda.is_current_stream_capturing():  # graphs: update_graphed() built the map already
                        self._set_downsample_map(m, keep)
                else:
                    L.downsample_map(sp.n_points, float(hp.aug_lo), int(hp.aug_hi != 0), self.seed, self.counter, sid, m, ST())
                maps[f"keep_{which}"] = m
            noise = dict(noise, **maps)
        do_actor = updates % hp.actor_update_interval == 0
        do_target = updates % hp.target_update_interval == 0
        ld_cat = D + S + A
        target_entropy = float(hp.target_entropy) if hp.target_entropy is not None else -float(A)

        # weight packing only feeds the encodes: it runs beside the two staging kernels
        s_p = self._fork(3)
        s_pb = self._fork(9)
        with torch.cuda.stream(s_pb):
            self._pack_weights(ST(), 2)  # read by the backward's recompute ~0.4 ms from now: off the critical path
        with torch.cuda.stream(s_p):
            self._pack_weights(ST(), 1)
            if self.actor_w0p is not None:  # refreshed every update: the parameters may have been written from outside
                L.copy_cols(p["actor.w0"], D + S, 1, 1, self.actor_w0p, self.actor_w0p.shape[1], 0, sp.hidden[0], D + S,
                            ST())

        # ---- branch T (side stream 0): TD target, no grad -- sac.py:108-134 / drq.py:71-87
        s_t = self._fork(0)
        with torch.cuda.stream(s_t):
            self._stage("next_obs", "next", k, aug, noise.get(f"{nkey}_next") if nkey else None, 1, ST())
            s_t.wait_stream(s_p)
            self._encode("next", R, False, ST())
            cat = w["cat_next"]
            if S:
                L.copy_cols(self.raw["next_obs"]["state"], S, k, 1, cat, ld_cat, D, R, S, ST())
            self._mlp_fwd("actor", cat, D + S, R, w["out_next"], 2 * A, 2 * A, "tmp", ST())
            L.tanh_gaussian_fwd(w["out_next"], R, A, hp.log_std_bound[0], hp.log_std_bound[1], hp.head_scale,
                                hp.head_bias, noise.get("eps_next"), self.seed, self.counter, 2, cat[:, D + S:], ld_cat,
                                w["nlp_next"], w["eps_next"], ST())
            s_t1 = self._fork(1)
            with torch.cuda.stream(s_t1):
                self._mlp_fwd("tq1", cat, ld_cat, R, w["q_next"][:, 1:], 2, 1, "tmp2", ST())
            self._mlp_fwd("tq0", cat, ld_cat, R, w["q_next"], 2, 1, "tmp", ST())
            self._join(s_t1)
            L.td_target(w["q_next"], w["nlp_next"], self.raw["rewards"], self.raw["dones"], B, k, hp.gamma,
                        1.0 if hp.algo == "drq" else hp.reward_scale, int(hp.ignore_dones), self.alpha_dev, w["y"], ST())

        # ---- critic forward on the main stream: sac.py:136 / drq.py:89
        self._stage("obs", "obs", k, aug, noise.get(f"{nkey}_obs") if nkey else None, 0, ST())
        self._join(s_p)
        self._encode("obs", R, True, ST())
        cat = w["cat_obs"]
        if S:
            L.copy_cols(self.raw["obs"]["state"], S, k, 1, cat, ld_cat, D, R, S, ST())
        L.copy_cols(self.raw["actions"], A, k, 1, cat, ld_cat, D + S, R, A, ST())
        c_lo, c_hi = self.layout.group_range["critic"]
        self.grads[c_lo:c_hi].zero_()
        s_q = self._fork(2)
        with torch.cuda.stream(s_q):
            self._mlp_fwd("q1", cat, ld_cat, R, w["q_obs"][:, 1:], 2, 1, "q1", ST())
        self._mlp_fwd("q0", cat, ld_cat, R, w["q_obs"], 2, 1, "q0", ST())
        self._join(s_q, s_t)
        L.critic_loss(w["q_obs"], w["y"], R, w["dq"], self.scalars, ST())

        # ---- critic backward: the two heads in parallel, their feature gradients add (sac.py:141-142)
        # the heads' weight gradients are deferred behind the dX chains (they would take SMs away from them) and joined
        # before Adam -- or, with a gradient all-reduce, before the Q heads' reduction, which is issued from a side
        # stream so that neither the weight gradients nor NCCL hold up the PointNet backward on the compute stream
        defer = True
        s_q = self._fork(2)
        with torch.cuda.stream(s_q):
            wg1 = self._mlp_bwd("q1", cat, ld_cat, R, w["dq"][:, 1:], 2, 1, "q1", w["dx1"], True, ST(), "b",
                                wstream=(3, 4, 5), defer=defer)
        wg0 = self._mlp_bwd("q0", cat, ld_cat, R, w["dq"], 2, 1, "q0", w["dx0"], True, ST(), "a", wstream=(6, 7, 8),
                            defer=defer)
        self._join(s_q)
        q_pending = None
        if self.allreduce is not None:
            # the two Q heads' gradients (97 % of the critic bytes) are final once their weight-gradient streams are:
            # reduce them on NCCL's stream while the PointNet head + sparse backward still run on the compute stream
            q_lo, q_hi = self.layout.q_range
            s_r = self._side[1]  # idle since the target branch joined
            for s_w in (*wg0, *wg1):
                s_r.wait_stream(s_w)
            with torch.cuda.stream(s_r):
                q_pending = self.allreduce(self.grads[c_lo + q_lo:c_lo + q_hi], async_op=True)
            wg0, wg1 = [s_r], []
        L.add_cols(w["dx0"], ld_cat, w["dx1"], ld_cat, w["dz"], D, R, D, ST())  # both heads' d/dfeature add up
        L.layernorm_bwd(w["dz"], D, w["xhat_obs"], w["rstd_obs"], p["pn.gf"], self.g["pn.gf"], self.g["pn.bef"],
                        w["dz"], R, D, ST())
        L.linear_bwd(w["pooled_obs"], c3, p["pn.wf"], w["dz"], D, self.g["pn.wf"], self.g["pn.bf"], w["dpooled"], c3,
                     None, 0, R, c3, D, self.tf32, ST())
        g = self.g
        self._join(s_pb)
        L.pointnet_bwd(w["xf_obs"], R, sp.n_points, sp.NP, sp.CP, sp.C, w["pooled_obs"], w["argmax_obs"], w["dpooled"],
                       p["pn.w0"], p["pn.b0"], p["pn.w1"], p["pn.g1"], p["pn.be1"], p["pn.w2"], p["pn.g2"],
                       p["pn.be2"], c1, c2, c3, sp.ln_eps, g["pn.w0"], g["pn.b0"], g["pn.w1"], g["pn.g1"], g["pn.be1"],
                       g["pn.w2"], g["pn.g2"], g["pn.be2"], w["scratch"], self.bwd_ws_bytes, self.tf32, w.get("xh_obs"),
                       w.get("wpack"), ST())
        if self.allreduce is not None:
            # the PointNet gradients (0.3 MB) are reduced on NCCL's stream while Adam already steps the two Q heads,
            # whose reduction has been in flight since their weight gradients finished; then Adam steps the PointNet
            s_r = self._side[1]
            s_r.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(s_r):
                pn_pending = self.allreduce(self.grads[c_lo:c_lo + self.layout.q_range[0]], async_op=True)
            q_pending.wait()
            self._join(*wg0, *wg1)
            self._adam_part("critic", self.layout.q_range, 0, hp.lr, hp.betas, 4, True, do_target, ST())
            pn_pending.wait()
            self._adam_part("critic", (0, self.layout.q_range[0]), 0, hp.lr, hp.betas, 4, False, False, ST())
        else:
            self._join(*wg0, *wg1)
            self._adam("critic", 0, hp.lr, hp.betas, 4, do_target, ST())  # + Polyak fused (sac.py:207-208)

        # ---- actor + alpha step: sac.py:161-205 / drq.py:114-155
        if do_actor:
            if k > 1:  # first augmentation of every sample (drq.py:115)
                if self.precision != "bf16":
                    self._take_first_aug(ST())  # the tensor-core path reads the obs staging in place (strided encode)
                name = "pi"
            else:
                name = "obs"
            self._pack_weights(ST(), 1)  # no backward follows this encode
            self._encode(name, B, False, ST())  # post-critic-step PointNet weights; output detached
            cat = w[f"cat_{name}"]
            if S:
                L.copy_cols(self.raw["obs"]["state"], S, 1, 1, cat, ld_cat, D, B, S, ST())
            self._mlp_fwd("actor", cat, D + S, B, w["out_pi"], 2 * A, 2 * A, "actor", ST())
            L.tanh_gaussian_fwd(w["out_pi"], B, A, hp.log_std_bound[0], hp.log_std_bound[1], hp.head_scale,
                                hp.head_bias, noise.get("eps_pi"), self.seed, self.counter, 3, cat[:, D + S:], ld_cat,
                                w["nlp_pi"], w["eps_pi"], ST())
            a_lo, a_hi = self.layout.group_range["actor"]
            self.grads[a_lo:a_hi].zero_()
            s_q = self._fork(2)
            with torch.cuda.stream(s_q):
                self._mlp_fwd("q1", cat, ld_cat, B, w["q_pi"][:, 1:], 2, 1, "q1", ST())
            self._mlp_fwd("q0", cat, ld_cat, B, w["q_pi"], 2, 1, "q0", ST())
            self._join(s_q)
            L.actor_loss(w["q_pi"], w["nlp_pi"], B, self.alpha_dev, p["log_alpha"], target_entropy, w["dq"],
                         w["dlog_alpha"], self.scalars, ST())
            s_q = self._fork(2)
            with torch.cuda.stream(s_q):
                self._mlp_bwd("q1", cat, ld_cat, B, w["dq"][:, 1:], 2, 1, "q1", w["dx1"], False, ST(), "b")
            self._mlp_bwd("q0", cat, ld_cat, B, w["dq"], 2, 1, "q0", w["dx0"], False, ST(), "a")
            self._join(s_q)
            da = w["dx0"][:, D + S:]
            L.add_cols(da, ld_cat, w["dx1"][:, D + S:], ld_cat, da, ld_cat, B, A, ST())
            L.tanh_gaussian_bwd_dev(w["out_pi"], w["eps_pi"], da, ld_cat, self.alpha_dev, B, A, hp.log_std_bound[0],
                                    hp.log_std_bound[1], hp.head_scale, w["dout"], ST())
            if self.allreduce is not None:
                # one all-reduce per layer, issued on NCCL's stream the moment that layer's weight gradient exists
                # (layer 2 first): only the last bucket -- layer 0, ~1 MB -- is left exposed after the backward
                wsides = self._mlp_bwd("actor", cat, D + S, B, w["dout"], 2 * A, 2 * A, "actor", None, True, ST(), "a",
                                       wstream=(3, 4, 5), join=False)
                al_hi = self.layout.group_range["alpha"][1]
                ent = self.layout.entries
                edges = [ent["actor.w0"][0], ent["actor.w1"][0], ent["actor.w2"][0], al_hi]  # (w, b) pairs; alpha rides with layer 2
                s_r = self._side[1]
                pend = []
                for layer in (2, 1, 0):
                    s_r.wait_stream(wsides[layer])
                    with torch.cuda.stream(s_r):
                        pend.append(self.allreduce(self.grads[edges[layer]:edges[layer + 1]], async_op=True))
                for h in pend:
                    h.wait()
                self._join(*wsides)
            else:
                self._mlp_bwd("actor", cat, D + S, B, w["dout"], 2 * A, 2 * A, "actor", None, True, ST(), "a", wstream=(3, 4, 5))
            self._adam("actor", 1, hp.actor_lr, hp.actor_betas, 8, False, ST())
            if hp.automatic_alpha_tuning:
                self._adam("alpha", 2, hp.alpha_lr, hp.alpha_betas, None, False, ST())
                self.refresh_alpha()
        self.counter.add_(1)

    @staticmethod
    def _set_downsample_map(m, keep):
        """Source map of RandomDownSample from the injected kept indices: dropped points read the first kept one."""
        keep = keep.to(device=m.device, dtype=torch.int64)
        m.copy_(keep[:1].to(torch.int32).expand_as(m))
        m[keep] = keep.to(torch.int32)

    # ------------------------------------------------------------------ CUDA graphs
    def update_graphed(self, updates: int, noise: Optional[Dict[str, torch.Tensor]] = None):
        """Same as update() but replayed from a CUDA graph: one launch per update instead of ~150.  Two graphs exist at
        most per (actor step?, target step?) combination.  `noise` (parity mode) is copied into static device buffers
        the captured kernels read, so the graph path can be compared with the reference draw for draw."""
        with torch.cuda.device(self.device):
            return self._update_graphed(updates, noise)

    def _update_graphed(self, updates, noise):
        hp = self.hp
        key = (updates % hp.actor_update_interval == 0, updates % hp.target_update_interval == 0,
               tuple(sorted(noise)) if noise else None)
        static = None
        if noise:
            static = self._noise_static.setdefault(key, {})
            for k_, v in noise.items():
                if k_.startswith("keep_"):  # variable-length index list: turned into the fixed-size source map here
                    self._set_downsample_map(self.w[f"ds_map_{k_[5:]}"], v)
                    static[k_] = v
                    continue
                if k_ not in static:
                    static[k_] = torch.empty_like(v, device=self.device)
                static[k_].copy_(v, non_blocking=True)
        g = self._graphs.get(key)
        if g is None:
            # warm-up outside capture (module loading, cudaFuncSetAttribute), on a side stream as torch requires
            state = (self.params.clone(), self.adam_m.clone(), self.adam_v.clone(), self.steps.clone(),
                     self.counter.clone(), self.alpha_dev.clone(), self.scalars.clone())
            side = torch.cuda.Stream(device=self.device)
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                self._update(updates, static)
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize(self.device)
            for dst, src in zip((self.params, self.adam_m, self.adam_v, self.steps, self.counter, self.alpha_dev,
                                 self.scalars), state):
                dst.copy_(src)
            g = torch.cuda.CUDAGraph()
            n0 = self.L.launches
            with torch.cuda.graph(g, stream=self._capture_stream):
                self._update(updates, static)
            self.graph_calls[key] = self.L.launches - n0  # C-ABI calls (>= 1 kernel each) replayed per launch
            # capture does not execute: state is untouched
            self._graphs[key] = g
        g.replay()
        return self.graph_calls[key]

    def close(self):
        """Drops the captured CUDA graphs (they hold NCCL kernels when a gradient all-reduce is attached: destroy them
        BEFORE the process group, or communicator teardown blocks)."""
        with torch.cuda.device(self.device):
            torch.cuda.synchronize(self.device)
            self._graphs.clear()
            self._noise_static.clear()
            torch.cuda.synchronize(self.device)

    # ------------------------------------------------------------------ pipelined host->device input path
    def make_pinned_batch(self, batch):
        """The replay sample laid out in pinned host memory (what a pinned replay ring would hand over): one contiguous
        buffer with the device layout; the returned dict holds its leaf views plus the buffer itself under "_flat"."""
        flat = torch.empty(self._batch_bytes, dtype=torch.uint8).pin_memory()
        views = self._batch_views(flat)
        for k, v in self._flatten_batch(batch).items():
            views[k].copy_(v)
        views["_flat"] = flat
        return views

    def to_device_batch(self, pinned):
        """A batch resident in HBM (for set_batch_device)."""
        return {"_flat": pinned["_flat"].to(self.device)}

    def set_batch_device(self, dev_batch):
        """Batch already resident in HBM: device-to-device adopt."""
        if "_flat" in dev_batch:
            self.raw_flat.copy_(dev_batch["_flat"], non_blocking=True)
            return
        for k, src in dev_batch.items():
            self._device_leaf(k).copy_(src, non_blocking=True)

    def _take_first_aug(self, st):
        """xf_pi[b] = xf_obs[b*k] (and the bf16 tile images): strided device copies, no kernel."""
        w, B, k = self.w, self.B, self.k
        w["xf_pi"].copy_(w["xf_obs"].view(B, k, *w["xf_obs"].shape[1:])[:, 0])
        if "xh_obs" in w:
            w["xh_pi"].view(B, -1).copy_(w["xh_obs"].view(B, k, -1)[:, 0])

    # ------------------------------------------------------------------ readback
    def read_scalars(self, updates: int, sync=True):
        """One device->host copy of everything update_parameters() logs (vs ~11 .item() syncs, sac.py:140-203)."""
        self.flush_scalars()
        if self.scalars_host is not None:
            with torch.cuda.device(self.device):
                self.scalars_host.copy_(self.scalars, non_blocking=True)
                if sync:
                    torch.cuda.current_stream().synchronize()
            s = self.scalars_host.numpy().astype(np.float64)
        else:
            s = self.scalars.cpu().numpy().astype(np.float64)
        return self._scalars_dict(s, updates)

    def _scalars_dict(self, s, updates):
        hp = self.hp
        pre = hp.algo
        A = self.spec.action_dim
        ret = {
            f"{pre}/critic_loss": float(s[0]),
            f"{pre}/max_critic_abs_err": float(s[1]),
            f"{pre}/alpha": float(self._alpha_before),
            f"{pre}/q": float(s[2]),
            f"{pre}/q_target": float(s[3]),
            f"{pre}/target_entropy": hp.target_entropy if hp.target_entropy is not None else -A,
            f"{pre}/critic_grad": float(np.sqrt(s[4])),
            f"{pre}/grad_steps": 1,
        }
        if updates % hp.actor_update_interval == 0:
            ret[f"{pre}/actor_loss"] = float(s[5])
            ret[f"{pre}/alpha_loss"] = float(s[6]) if hp.automatic_alpha_tuning else 0.0
            ret[f"{pre}/entropy"] = float(s[7])
            ret[f"{pre}/actor_grad"] = float(np.sqrt(s[8]))
        self._alpha_before = float(s[9])
        return ret

    def read_scalars_async(self, updates: int):
        """The same dict, filled in on first access: the device->host copy of the scalars is enqueued behind the update
        (pinned ring slot + event) and this call returns at once, so the host can sample and stage the NEXT batch while
        this update runs.  Reading any entry waits for the event.  Pending results resolve in update order (the alpha an
        update logs is the one the update before it left, sac.py:152); a slot is never reused before its result was
        resolved."""
        if self.scalars_host is None:
            return self.read_scalars(updates)
        if self._scalar_ring is None:
            self._scalar_ring = [torch.zeros(NUM_SCALARS, dtype=torch.float32).pin_memory() for _ in range(SCALAR_RING)]
            self._scalar_pending = []
            self._scalar_seq = 0
        while len(self._scalar_pending) >= SCALAR_RING:
            self._scalar_pending[0]._resolve()
        slot = self._scalar_ring[self._scalar_seq % SCALAR_RING]
        self._scalar_seq += 1
        with torch.cuda.device(self.device):
            slot.copy_(self.scalars, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record()
        ret = LazyScalars(self, slot, ev, updates)
        self._scalar_pending.append(ret)
        return ret

    def flush_scalars(self):
        """Resolve every pending asynchronous result (keeps `_alpha_before` and the ring in update order)."""
        while self._scalar_pending:
            self._scalar_pending[0]._resolve()

    _scalar_ring = None
    _scalar_pending = ()
    _alpha_before = 0.0

    def prime_alpha(self):
        """host copy of the cached alpha (the value update N logs is the one cached BEFORE update N, sac.py:152)."""
        self.flush_scalars()
        self._alpha_before = float(self.alpha_dev.item())
