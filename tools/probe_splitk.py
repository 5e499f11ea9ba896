"""Time pcrl_gemm_tf32 for the MLP-head shapes with different split-K factors (atomic accumulation), warm, in-graph."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pointcloud_rl_b200._lib import lib, stream_ptr
from tools.probe_mlp import timed_graph  # noqa
L = lib()
dev = "cuda"
for (M, N, K) in [(512, 1024, 1024), (512, 1024, 256), (256, 1024, 1024), (512, 44, 1024), (512, 256, 1024)]:
    x = torch.randn(M, K, device=dev); w = torch.randn(N, K, device=dev) * 0.03
    y = torch.zeros(M, N, device=dev)
    for sk in (1, 2, 4, 8):
        mode = 0 if sk == 1 else 2
        t = timed_graph(lambda: L.gemm_tf32(x, K, 0, w, K, 0, None, y, N, M, N, K, 0, mode, sk, stream_ptr()))
        print(f"M={M} N={N} K={K} split_k={sk}: {t:7.2f} us")
