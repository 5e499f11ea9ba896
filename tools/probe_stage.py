"""CUDA-event timing of the staging kernel alone at the benchmark shapes (L2 flushed between launches)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from bench import WORKLOADS, build_engine
from pointcloud_rl_b200._lib import stream_ptr
from pointcloud_rl_b200.synthetic import synthetic_batch

name = sys.argv[1] if len(sys.argv) > 1 else "drq_maniskill_pn_jitter"
w = WORKLOADS[name]
eng, spec = build_engine(w, "bf16", "cuda:0", 0)
b = synthetic_batch(0, w["B"], w["N"], w["A"], n_seg=w["n_seg"], n_pos=w["n_pos"], state_dim=w["S"])
eng.upload_batch(b)
st = stream_ptr()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
aug = 1 if w["aug"] == "jitter" else 0
ts = []
for i in range(25):
    flush.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); eng._stage("next_obs", "next", eng.k, aug, None, 1, st); e1.record()
    torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1) * 1e3)
rows = eng.R * spec.NP
print(f"{name}: stage {eng.R} clouds x {spec.NP} rows: {np.median(ts[5:]):.1f} us median, {min(ts[5:]):.1f} min "
      f"(variant {os.environ.get('PCRL_STAGE_VARIANT', '0')})")
