"""Debug driver for the wide (64-128-1024) workload: python tools/wide_dbg.py B N [graph]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import WORKLOADS, build_engine
from pointcloud_rl_b200.synthetic import synthetic_batch
B, N = int(sys.argv[1]), int(sys.argv[2])
graph = len(sys.argv) > 3 and sys.argv[3] == "graph"
w = dict(WORKLOADS["sac_wide"], B=B, N=N)
eng, spec = build_engine(w, "bf16", "cuda:0", 0)
eng.upload_batch(synthetic_batch(0, B, N, w["A"]))
for u in (1, 2, 3, 4):
    (eng.update_graphed if graph else eng.update)(u)
    torch.cuda.synchronize()
    print("update", u, "ok", {k: round(v, 4) for k, v in eng.read_scalars(u).items() if "loss" in k})
