"""Data-parallel plumbing: one process per GPU, NCCL all-reduce of the flat gradient buffers.

Replaces the reference's ExtendedDDP wrapping (module_utils.py:105-110,322-349): instead of DDP's bucketed
hooks, the engine all-reduces its two flat gradient ranges (critic+PointNet after the critic backward;
actor MLP + d log_alpha after the actor backward) and folds 1/world into the fused Adam kernel.
Deliberate deviation: log_alpha's gradient is reduced too, so alpha stays identical on every rank (the
reference leaves it un-synced, sac.py:83 + module_utils.py:338-343).
"""
import os

import torch
import torch.distributed as dist

from ._lib import stream_ptr


class _EventHandle:
    """What `allreduce(..., async_op=True)` returns on the peer-memory path: `wait()` makes the current stream wait for
    the reduction kernel (an event dependency -- also inside CUDA-graph capture), like a c10d work handle does."""

    def __init__(self, event):
        self.event = event

    def wait(self):
        torch.cuda.current_stream().wait_event(self.event)


class PeerAllReduce:
    """In-place SUM all-reduce of ranges of the engine's flat gradient buffer through NVLink peer memory
    (`pcrl_p2p_allreduce`, csrc/p2p.cu): the gradient buffer is re-allocated as symmetric memory (same layout on every
    rank of the node, mapped into every peer), one kernel per reduction, bit-identical result on all ranks."""

    def __init__(self, engine, group):
        import torch.distributed._symmetric_memory as symm

        if engine._graphs:
            raise RuntimeError("the gradient buffer cannot move once update graphs were captured")
        L, dev = engine.L, engine.device
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        name = group.group_name
        grads = symm.empty(engine.grads.numel(), dtype=torch.float32, device=dev)
        self.h_grads = symm.rendezvous(grads, name)
        flags = symm.empty(int(L.p2p_flag_bytes()) // 4, dtype=torch.int32, device=dev)
        self.h_flags = symm.rendezvous(flags, name)
        grads.zero_()
        flags.zero_()
        torch.cuda.synchronize(dev)
        self.h_flags.barrier()  # nobody signals into a flag block that is not zeroed yet
        self.grads, self.flags = grads, flags
        self.bufs_dev = torch.tensor([int(p) for p in self.h_grads.buffer_ptrs], dtype=torch.int64, device=dev)
        self.flags_dev = torch.tensor([int(p) for p in self.h_flags.buffer_ptrs], dtype=torch.int64, device=dev)
        self.state = torch.zeros(int(L.p2p_state_bytes()) // 4, dtype=torch.int32, device=dev)
        self.channels = {}
        self.L = L
        engine.rebind_grads(grads)

    def __call__(self, flat_grad, async_op=False):
        if flat_grad.untyped_storage().data_ptr() != self.grads.untyped_storage().data_ptr() or not flat_grad.is_contiguous():
            raise RuntimeError("peer-memory all-reduce: not a contiguous range of the engine's gradient buffer")
        key = (flat_grad.storage_offset(), flat_grad.numel())
        ch = self.channels.setdefault(key, len(self.channels))  # one channel per call site, same order on every rank
        self.L.p2p_allreduce(self.bufs_dev, self.flags_dev, self.rank, self.world, key[0], key[1], ch, self.state, 0,
                             stream_ptr())
        if not async_op:
            return None
        ev = torch.cuda.Event()
        ev.record()
        return _EventHandle(ev)

    def check(self):
        """Raises if a reduction gave up waiting for a peer (state[channel][2] != 0)."""
        err = self.state.view(-1, 4)[:, 2].cpu()
        if int(err.abs().sum()):
            raise RuntimeError(f"peer-memory all-reduce timed out waiting for a peer (per channel: {err.tolist()})")


def attach(engine, group=None, peer_memory=None):
    """Installs the gradient all-reduce.  peer_memory: False = NCCL; True = the NVLink peer-memory kernel (all ranks on
    one node); None = NCCL unless PCRL_P2P_ALLREDUCE=1 is set.  NCCL is the default: the peer-memory kernel is validated
    on its own at 2 and 8 ranks and in the engine at 2 ranks (tests/test_gpu_nccl.py), but the 8-rank benchmark run with
    it did not finish this round, so it stays opt-in.  When it is requested the ranks agree on the
    outcome: a rank whose setup fails takes everybody back to NCCL."""
    group = group if group is not None else dist.group.WORLD
    world = dist.get_world_size(group)
    engine.world_size = world

    def allreduce(flat_grad: torch.Tensor, async_op: bool = False):
        """SUM over ranks (1/world is folded into the fused Adam).  async_op=True returns a work handle: NCCL runs on
        its own stream and `handle.wait()` re-joins the compute stream, so the reduce overlaps later kernels."""
        return dist.all_reduce(flat_grad, op=dist.ReduceOp.SUM, group=group, async_op=async_op)

    engine.allreduce = allreduce if world > 1 else None
    engine.allreduce_kind = "nccl" if world > 1 else None
    if peer_memory is None:
        peer_memory = os.environ.get("PCRL_P2P_ALLREDUCE", "0") == "1"
    device = getattr(engine, "device", None)
    if not peer_memory or world == 1 or device is None or device.type != "cuda" or dist.get_backend(group) != "nccl":
        return engine  # the NCCL path: no further collective here
    want = world <= 16 and not engine._graphs
    ok = torch.tensor([1 if want else 0], device=device)
    dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)
    if not int(ok.item()):
        return engine
    p2p, err = None, None
    try:
        p2p = PeerAllReduce(engine, group)
    except Exception as e:  # noqa: BLE001 -- symmetric memory unavailable (ranks on several nodes, no P2P, ...)
        err = e
    ok = torch.tensor([0 if p2p is None else 1], device=device)
    dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)
    if int(ok.item()):
        engine.allreduce, engine.allreduce_kind, engine._p2p = p2p, "peer_memory", p2p
    elif err is not None:
        import warnings

        warnings.warn(f"peer-memory all-reduce unavailable, using NCCL: {err}")
    return engine


def broadcast_params(engine, src=0, group=None):
    """Make every rank start from rank `src`'s weights (what DDP does at wrap time)."""
    dist.broadcast(engine.params, src=src, group=group)
    engine.refresh_alpha()


def broadcast_state(engine, src=0, group=None):
    """Parameters, target networks, Adam moments and step counters of rank `src` on every rank: the replicas then stay
    identical because they all apply the same reduced gradient."""
    for t in (engine.params, engine.adam_m, engine.adam_v, engine.steps):
        dist.broadcast(t, src=src, group=group)
    engine.refresh_alpha()
    engine.prime_alpha()
