"""Time the MLP-head GEMM chains of one update (warm L2, replayed from a CUDA graph) next to cuBLAS TF32."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pointcloud_rl_b200._lib import lib, stream_ptr

L = lib()
torch.backends.cuda.matmul.allow_tf32 = True
dev = "cuda"

def timed_graph(fn, reps=20, iters=10):
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        fn(); torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            for _ in range(reps):
                fn()
        g.replay(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(s)
        for _ in range(iters):
            g.replay()
        e1.record(s); torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / (reps * iters)

if __name__ == '__main__':
  for M in (512, 256):
      for (K, N, relu) in [(256, 1024, 1), (1024, 1024, 1), (1024, 44, 0), (1024, 1, 0)]:
          x = torch.randn(M, K, device=dev); w = torch.randn(N, K, device=dev) * 0.03; b = torch.randn(N, device=dev)
          y = torch.empty(M, N, device=dev)
          t = timed_graph(lambda: L.linear_fwd(x, K, w, b, y, N, M, K, N, relu, 1, stream_ptr()))
          tb = timed_graph(lambda: torch.addmm(b, x, w.t(), out=y))
          print(f"fwd  M={M:4d} K={K:5d} N={N:5d}: pcrl {t:7.2f} us   cuBLAS(tf32 addmm) {tb:7.2f} us")
      for (K, N) in [(1024, 1024), (256, 1024), (1024, 44)]:
          # y = x W^T (x [M,K], W [N,K]); backward: dW [N,K] = dy^T x, dx [M,K] = dy W
          x = torch.randn(M, K, device=dev); w = torch.randn(N, K, device=dev) * 0.03
          dy = torch.randn(M, N, device=dev); dw = torch.zeros(N, K, device=dev); db = torch.zeros(N, device=dev)
          dx = torch.empty(M, K, device=dev); mask = torch.randn(M, K, device=dev)
          t = timed_graph(lambda: L.linear_bwd(x, K, w, dy, N, dw, db, dx, K, mask, K, M, K, N, 1, stream_ptr()))
          t2 = timed_graph(lambda: L.linear_bwd(x, K, w, dy, N, None, None, dx, K, mask, K, M, K, N, 1, stream_ptr()))
          tb = timed_graph(lambda: (torch.mm(dy.t(), x, out=dw), torch.mm(dy, w, out=dx)))
          print(f"bwd  M={M:4d} K={K:5d} N={N:5d}: pcrl dW+db+dX {t:7.2f} us, dX only {t2:7.2f} us   cuBLAS dW+dX {tb:7.2f} us")
