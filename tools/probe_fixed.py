"""Fixed cost of one tc_gemm launch (tiny K) vs K, warm, replayed from a CUDA graph."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pointcloud_rl_b200._lib import lib, stream_ptr
from tools.probe_mlp import timed_graph
L = lib(); dev = "cuda"
torch.backends.cuda.matmul.allow_tf32 = True
for (M, N) in [(512, 1024), (512, 128)]:
    for K in (32, 64, 128, 256, 512, 1024):
        x = torch.randn(M, K, device=dev); w = torch.randn(N, K, device=dev) * 0.03; b = torch.randn(N, device=dev)
        y = torch.empty(M, N, device=dev)
        t = timed_graph(lambda: L.linear_fwd(x, K, w, b, y, N, M, K, N, 1, 1, stream_ptr()))
        tb = timed_graph(lambda: torch.addmm(b, x, w.t(), out=y))
        print(f"M={M} N={N} K={K:5d}: pcrl {t:6.2f} us   cuBLAS {tb:6.2f} us")
def empty():
    torch.cuda._sleep(0) if False else None
