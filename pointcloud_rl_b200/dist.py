"""Data-parallel plumbing: one process per GPU, NCCL all-reduce of the flat gradient buffers.

Replaces the reference's ExtendedDDP wrapping (module_utils.py:105-110,322-349): instead of DDP's bucketed
hooks, the engine all-reduces its two flat gradient ranges (critic+PointNet after the critic backward;
actor MLP + d log_alpha after the actor backward) and folds 1/world into the fused Adam kernel.
Deliberate deviation: log_alpha's gradient is reduced too, so alpha stays identical on every rank (the
reference leaves it un-synced, sac.py:83 + module_utils.py:338-343).
"""
import torch
import torch.distributed as dist


def attach(engine, group=None):
    group = group if group is not None else dist.group.WORLD
    world = dist.get_world_size(group)
    engine.world_size = world

    def allreduce(flat_grad: torch.Tensor, async_op: bool = False):
        """SUM over ranks (1/world is folded into the fused Adam).  async_op=True returns a work handle: NCCL runs on
        its own stream and `handle.wait()` re-joins the compute stream, so the reduce overlaps later kernels."""
        return dist.all_reduce(flat_grad, op=dist.ReduceOp.SUM, group=group, async_op=async_op)

    engine.allreduce = allreduce if world > 1 else None
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
