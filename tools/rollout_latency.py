"""Rollout-side latency (BaseAgent.forward: obs -> device -> PointNet encode -> actor MLP -> tanh-Gaussian head) for the
ManiSkill DrQ agent at B = num_envs = 1, 4, 16, 64: wall-clock per call including the host->device copy of the
observation and the device->host read of the action, cached weight images (no update between calls)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from bench import WORKLOADS, build_bench_agent
from pointcloud_rl_b200.synthetic import synthetic_obs

w = WORKLOADS["drq_maniskill_pn_jitter"]
print("# B, precision, us per agent(obs) call (median of 200), of which device-side (CUDA events)")
for prec in ("bf16", "tf32", "fp32"):
    agent = build_bench_agent(w, prec, "cuda:0", 0)
    for B in (1, 4, 16, 64):
        obs = synthetic_obs(np.random.RandomState(B), B, w["N"], n_seg=w["n_seg"], state_dim=w["S"])
        for _ in range(10):
            agent(obs, mode="explore").cpu()
        wall, dev = [], []
        for _ in range(200):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0 = time.perf_counter()
            e0.record()
            a = agent(obs, mode="explore")
            e1.record()
            a = a.cpu()
            wall.append((time.perf_counter() - t0) * 1e6)
            dev.append(e0.elapsed_time(e1) * 1e3)
        print(f"{B}, {prec}, {np.median(wall):.1f}, {np.median(dev):.1f}")
