"""BASELINE config 3: PointNet encoder forward-only (rollout/actor path) throughput sweep, B = 1..4096, N = 1200,
fp32 (exact FFMA path) vs the bf16 tcgen05 path.  Points/s counts real points; L2 is flushed between launches.
Usage (GPU box): python tools/encoder_sweep.py > profiles/rNN_encoder_sweep.txt"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from pointcloud_rl_b200.engine import PathSpec
from pointcloud_rl_b200.networks import KernelRunner
from pointcloud_rl_b200.synthetic import init_params, synthetic_obs

N = 1200
spec = PathSpec(n_points=N, action_dim=22, state_dim=106, n_seg=1, widths=(128, 128, 256), out_dim=128)
p = {k: v.cuda() for k, v in init_params(0, spec).items()}
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
F_pt = 2 * (spec.C * 128 + 128 * 128 + 128 * 256)
print(f"# PointNet encode (stage + per-point MLP + max-pool + final Linear/LN), N={N}, C={spec.C}, widths 128-128-256; F_pt={F_pt} FLOP/point")
print("# B, precision, ms, Mpoints/s, TFLOP/s")
for prec in ("bf16", "fp32"):
    for B in (1, 2, 4, 8, 16, 32, 64, 128, 256, 512, 1024, 2048, 4096):
        if prec == "fp32" and B > 1024:
            continue
        rs = np.random.RandomState(B)
        obs = {k: torch.from_numpy(v).cuda() for k, v in synthetic_obs(rs, B, N, n_seg=1).items()}
        obs["seg"] = obs["seg"].to(torch.uint8)
        run = KernelRunner(prec)
        for _ in range(3):
            run.encode(spec, p, obs)
        ts = []
        for _ in range(8):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); run.encode(spec, p, obs); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ms = float(np.median(ts))
        print(f"{B}, {prec}, {ms:.4f}, {B * N / ms / 1e3:.1f}, {B * N * F_pt / ms / 1e9:.1f}")
