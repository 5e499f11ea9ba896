"""Data-parallel host logic on CPU: world_size-2 gloo run of the gradient all-reduce hook (dist.attach)
and of the 'sum then scale by 1/world inside Adam' convention the engine uses."""
import os

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


class _FakeEngine:
    """The attributes dist.attach touches on UpdateEngine."""

    def __init__(self):
        self.world_size = 1
        self.allreduce = None
        self.params = torch.zeros(8)
        self.adam_m, self.adam_v = torch.zeros(8), torch.zeros(8)
        self.steps = torch.zeros(4, dtype=torch.int32)
        self.primed = 0

    def refresh_alpha(self):
        pass

    def prime_alpha(self):
        self.primed += 1


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from pointcloud_rl_b200.dist import attach, broadcast_params, broadcast_state

    eng = attach(_FakeEngine())
    assert eng.world_size == world and eng.allreduce is not None and eng.allreduce_kind == "nccl"
    # the NVLink peer-memory kernel needs CUDA + NCCL: asking for it on a CPU / gloo group quietly keeps the plain
    # all-reduce (and must not issue a collective the other rank does not)
    eng2 = attach(_FakeEngine(), peer_memory=(rank == 0))
    assert eng2.allreduce_kind == "nccl" and eng2.allreduce is not None
    # each rank's gradient is a mean over its own equal-sized shard; sum * (1/world) == global mean
    torch.manual_seed(0)
    per_sample = torch.randn(world * 4, 6)
    local = per_sample[rank * 4:(rank + 1) * 4].mean(0)
    flat = local.clone()
    eng.allreduce(flat)
    scaled = flat * (1.0 / eng.world_size)
    ok = torch.allclose(scaled, per_sample.mean(0), atol=1e-6)
    eng.params = torch.full((8,), float(rank + 1))
    broadcast_params(eng, src=0)
    ok = ok and bool((eng.params == 1.0).all())
    # to_ddp(): parameters, Adam moments and step counters all follow rank 0 (what DDP wrapping does for the weights)
    eng.params = torch.full((8,), float(rank + 3))
    eng.adam_m = torch.full((8,), float(rank + 5))
    eng.adam_v = torch.full((8,), float(rank + 7))
    eng.steps = torch.full((4,), rank + 2, dtype=torch.int32)
    broadcast_state(eng, src=0)
    ok = ok and bool((eng.params == 3.0).all() and (eng.adam_m == 5.0).all() and (eng.adam_v == 7.0).all())
    ok = ok and bool((eng.steps == 2).all()) and eng.primed == 1
    if rank == 0:
        out.put(ok)
    dist.barrier()
    dist.destroy_process_group()


def test_gradient_allreduce_world2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 500
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert q.get(timeout=5) is True
