"""Kernel-by-kernel timeline of the TF32-tier PointNet forward chain (pcrl_pointnet_fwd_tf32) on 512 clouds x 1200 points."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import ProfilerActivity, profile
from bench import WORKLOADS, build_engine
from pointcloud_rl_b200._lib import stream_ptr
from pointcloud_rl_b200.synthetic import synthetic_batch

w = WORKLOADS["drq_maniskill_pn_jitter"]
eng, spec = build_engine(w, sys.argv[1] if len(sys.argv) > 1 else "tf32", "cuda:0", 0)
eng.upload_batch(synthetic_batch(0, w["B"], w["N"], w["A"], n_seg=w["n_seg"], state_dim=w["S"]))
st = stream_ptr()
eng._stage("next_obs", "next", eng.k, 1, None, 1, st)
for _ in range(3):
    eng._encode_points("next", eng.R, False, st)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    eng._encode_points("next", eng.R, False, st)
    torch.cuda.synchronize()
evs = sorted([e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA], key=lambda e: e.time_range.start)
t0 = evs[0].time_range.start
tot = {}
for e in evs[:14]:
    print(f"{e.time_range.start - t0:9.1f} {e.time_range.end - e.time_range.start:7.1f}  {e.name[:90]}")
for e in evs:
    k = e.name.split("(")[0][-60:]
    tot[k] = tot.get(k, 0) + (e.time_range.end - e.time_range.start)
print("totals (us):", {k: round(v, 1) for k, v in sorted(tot.items(), key=lambda kv: -kv[1])})
print("span us:", evs[-1].time_range.end - t0)
