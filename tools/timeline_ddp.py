"""Kernel timeline (CUPTI) of two graph-replayed data-parallel updates on rank 0 (torchrun, >= 2 GPUs): shows where the
NCCL all-reduces sit relative to the compute chain.  Usage: torchrun --nproc-per-node N tools/timeline_ddp.py > out.txt"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
from torch.profiler import ProfilerActivity, profile
from bench import WORKLOADS, build_engine
from pointcloud_rl_b200.dist import attach, broadcast_state
from pointcloud_rl_b200.synthetic import synthetic_batch

rank, lr = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device(f"cuda:{lr}"))
w = WORKLOADS["drq_maniskill_pn_jitter"]
eng, spec = build_engine(w, "bf16", f"cuda:{lr}", 1234 + rank)
attach(eng)
broadcast_state(eng)
eng.upload_batch(synthetic_batch(100 * rank, w["B"], w["N"], w["A"], n_seg=w["n_seg"], state_dim=w["S"]))
for u in range(1, 13):
    eng.update_graphed(u)
torch.cuda.synchronize(); dist.barrier()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for u in range(13, 15):
        eng.update_graphed(u)
    torch.cuda.synchronize()
if rank == 0:
    evs = sorted([e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA], key=lambda e: e.time_range.start)
    t0 = evs[0].time_range.start
    print(f"# world {dist.get_world_size()}  start_us  dur_us  kernel")
    for e in evs:
        name = e.name.split("(")[0].replace("void ", "").split("::")[-1][:50]
        mark = "  <<< NCCL" if "nccl" in e.name.lower() else ""
        print(f"{e.time_range.start - t0:9.1f} {e.time_range.end - e.time_range.start:7.1f}  {name}{mark}")
    print("# span us:", evs[-1].time_range.end - t0)
eng.close()
torch.cuda.synchronize(); dist.barrier(); dist.destroy_process_group()
