"""Worker of tests/test_gpu_nccl.py (launched by torch.distributed.run, one process per GPU).

Every rank takes its contiguous shard of ONE golden batch and of the injected noise, all-reduces its gradients over NCCL
(captured inside the CUDA graph when --graph), and must end up with (a) bit-identical parameters / Adam state on all ranks
and (b) the parameters a single engine computes on the whole batch (mean-of-shard-means == global mean for equal shards).
Exit code 0 = all checks passed on every rank.
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from tests.conftest import load_golden  # noqa: E402


def shard(tree, rank, world):
    if isinstance(tree, dict):
        return {k: shard(v, rank, world) for k, v in tree.items()}
    n = tree.shape[0] // world
    return tree[rank * n:(rank + 1) * n]


def main():
    precision, use_graph = sys.argv[1], sys.argv[2] == "graph"
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = f"cuda:{local}"
    dist.init_process_group("nccl", device_id=torch.device(dev))
    from pointcloud_rl_b200.dist import attach, broadcast_state
    from pointcloud_rl_b200.engine import HyperParams, PathSpec, UpdateEngine

    g = load_golden("drq_jitter_small")
    m = {k: v.item() for k, v in g["meta"].items()}
    init = {k: torch.from_numpy(np.asarray(v)) for k, v in g["init"].items()}
    c1, c2, c3 = init["pn.w0"].shape[0], init["pn.w1"].shape[0], init["pn.w2"].shape[0]
    spec = PathSpec(n_points=m["N"], action_dim=m["A"], state_dim=m["S"], n_pos=m["n_pos"], n_seg=m["n_seg"],
                    widths=(c1, c2, c3), out_dim=init["pn.wf"].shape[0],
                    hidden=(init["actor.w0"].shape[0], init["actor.w1"].shape[0]))
    hp = HyperParams(algo=m["algo"], gamma=m["gamma"], num_aug=m["num_aug"], aug=m["aug"], aug_lo=m["aug_lo"],
                     aug_hi=m["aug_hi"], tau=m["tau"], target_entropy=m["target_entropy"])
    B = m["B"]
    assert B % world == 0
    ok = True

    whole = UpdateEngine(spec, hp, batch_size=B, device=dev, precision=precision)  # the single-process answer
    whole.load_params(init)
    whole.prime_alpha()
    whole.upload_batch(g["batch"])
    part = UpdateEngine(spec, hp, batch_size=B // world, device=dev, precision=precision)
    if rank == 0:
        part.load_params(init)
    else:  # other ranks start from garbage: broadcast_state must bring rank 0's weights over
        part.load_params({k: torch.randn_like(v) for k, v in init.items()})
    reduce_kind = sys.argv[3] if len(sys.argv) > 3 else "nccl"
    attach(part, peer_memory=(reduce_kind == "peer_memory"))
    assert part.allreduce_kind == reduce_kind, (part.allreduce_kind, reduce_kind)
    broadcast_state(part)
    part.upload_batch(shard(g["batch"], rank, world))
    for u in range(1, m["n_updates"] + 1):
        # Every update starts, on all ranks and in both engines, from rank 0's single-process state (what to_ddp()'s
        # broadcast does before training).  Otherwise the ranks' single-process engines differ in the last bits (float
        # atomics) and, in the reduced-precision tier, trajectories drift apart through Adam (|delta w| ~ lr whatever
        # the gradient's size) -- neither says anything about the all-reduce.
        for t in (whole.params, whole.adam_m, whole.adam_v, whole.steps, whole.counter):
            dist.broadcast(t, src=0)
        whole.refresh_alpha()
        whole.prime_alpha()
        for dst, src in ((part.params, whole.params), (part.adam_m, whole.adam_m), (part.adam_v, whole.adam_v),
                         (part.steps, whole.steps), (part.counter, whole.counter)):
            dst.copy_(src)
        part.refresh_alpha()
        part.prime_alpha()
        noise = {k: torch.from_numpy(np.asarray(v)) for k, v in g[f"noise{u}"].items()}
        whole.update(u, {k: v.to(dev) for k, v in noise.items()})
        mine = {k: v.to(dev) for k, v in shard(noise, rank, world).items()}
        if use_graph:
            part.update_graphed(u, mine)
        else:
            part.update(u, mine)
        torch.cuda.synchronize()
        state = torch.cat([part.params.flatten(), part.adam_m.flatten(), part.adam_v.flatten()])
        ref = state.clone()
        dist.broadcast(ref, src=0)
        same = torch.equal(state, ref)
        # vs the whole-batch engine: Adam turns a rounding-level gradient difference into at most 2*lr per element
        # (sign flips of near-zero gradients), so judge the vector norm-wise and the gradient itself tightly
        dp = float((part.params - whole.params).norm() / whole.params.norm())
        lo, hi = part.layout.group_range["critic"]
        gsum = part.grads[lo:hi] / world
        dg = float((gsum - whole.grads[lo:hi]).norm() / whole.grads[lo:hi].norm())
        tol_g = 1e-4 if precision == "fp32" else 1e-3  # same kernels on the same clouds: only summation order differs
        good = same and dg < tol_g and dp < (1e-4 if precision == "fp32" else 1e-3)
        if not good:
            print(f"[rank {rank}] update {u}: identical={same} grad_rel={dg:.2e} param_rel={dp:.2e}", flush=True)
        ok = ok and good
    flag = torch.tensor([int(ok)], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("NCCL_2RANK", "PASS" if flag.item() == 1 else "FAIL", precision, "graph" if use_graph else "eager",
              part.allreduce_kind, flush=True)
    if getattr(part, "_p2p", None) is not None:
        part._p2p.check()
    # teardown order that does not hang: captured graphs hold NCCL kernels, so they go before the communicator
    part.close()
    whole.close()
    torch.cuda.synchronize()
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if flag.item() == 1 else 1)


if __name__ == "__main__":
    main()
