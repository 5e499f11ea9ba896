"""2+ rank consistency check (run under torchrun on a GPU box): every rank trains on its own batches and Philox stream, the
flat gradient ranges are all-reduced inside the captured graphs, so after several updates all ranks must hold bit-identical
parameters and Adam state, and they must differ from a rank trained alone (the reduction really happened)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
from bench import WORKLOADS, build_engine
from pointcloud_rl_b200.dist import attach
from pointcloud_rl_b200.synthetic import synthetic_batch

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device(f"cuda:{lr}"))
w = dict(WORKLOADS["drq_maniskill_pn_jitter"]); w["B"] = 32
eng, spec = build_engine(w, "bf16", f"cuda:{lr}", seed=1234 + rank)
attach(eng, dist.group.WORLD)
solo, _ = build_engine(w, "bf16", f"cuda:{lr}", seed=1234 + rank)
for u in range(1, 7):
    b = synthetic_batch(100 * rank + u, w["B"], w["N"], w["A"], n_seg=w["n_seg"], state_dim=w["S"])
    eng.upload_batch(b); eng.update_graphed(u)
    solo.upload_batch(b); solo.update_graphed(u)
torch.cuda.synchronize()
mine = torch.cat([eng.params.double().flatten(), eng.adam_m.double().flatten(), eng.adam_v.double().flatten()])
ref = mine.clone(); dist.broadcast(ref, src=0)
same = bool(torch.equal(mine, ref))
diff_solo = float((eng.params - solo.params).abs().max())
flags = torch.tensor([int(same)], device=f"cuda:{lr}"); dist.all_reduce(flags, op=dist.ReduceOp.MIN)
if rank == 0:
    print(f"world={world}: params/Adam state identical on all ranks: {bool(flags.item())}; max |param - solo-trained param| = {diff_solo:.3e}")
torch.cuda.synchronize()
sys.stdout.flush()
os._exit(0 if flags.item() == 1 and diff_solo > 0 else 1)
