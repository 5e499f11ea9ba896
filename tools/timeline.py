"""Kernel timeline of one graph-replayed update (CUPTI through torch.profiler): start, duration, stream of every kernel.
Usage (GPU box): python tools/timeline.py [n_updates] > gpurun_out/timeline.txt"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import ProfilerActivity, profile
from bench import WORKLOADS, build_engine
from pointcloud_rl_b200.synthetic import synthetic_batch

w = WORKLOADS["drq_maniskill_pn_jitter"]
eng, spec = build_engine(w, "bf16", "cuda:0", 0)
eng.upload_batch(synthetic_batch(0, w["B"], w["N"], w["A"], n_seg=w["n_seg"], state_dim=w["S"]))
for u in range(1, 9):
    eng.update_graphed(u)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for u in range(9, 11):
        eng.update_graphed(u)
    torch.cuda.synchronize()
evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
evs.sort(key=lambda e: e.time_range.start)
t0 = evs[0].time_range.start
streams = {}
print("# start_us  dur_us  stream  kernel")
for e in evs:
    sid = streams.setdefault(e.device_index if False else getattr(e, "stream", None) or 0, len(streams))
    name = e.name.split("(")[0].replace("void ", "").split("::")[-1][:44]
    print(f"{e.time_range.start - t0:9.1f} {e.time_range.end - e.time_range.start:7.1f}  s{sid}  {name}")
