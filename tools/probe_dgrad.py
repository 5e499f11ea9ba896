"""PointNet-backward-shaped dgrad GEMM (100 k rows): dX[M,128] = dY[M,256] W[256,128] with the fused ReLU mask, L2 flushed."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pointcloud_rl_b200._lib import lib, stream_ptr
L = lib(); dev = "cuda"
M, K, N = 100000, 128, 256
x = torch.randn(M, K, device=dev); w = torch.randn(N, K, device=dev) * 0.05
dy = torch.randn(M, N, device=dev); dx = torch.empty(M, K, device=dev); mask = torch.randn(M, K, device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for use_mask in (True, False):
    ts = []
    for i in range(8):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        L.linear_bwd(x, K, w, dy, N, None, None, dx, K, mask if use_mask else None, K if use_mask else 0, M, K, N, 1, stream_ptr())
        e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1) * 1e3)
    mb = (M * N * 4 + M * K * 4 * (2 if use_mask else 1)) / 1e6
    t = min(ts[2:])
    print(f"mask={use_mask}: {t:6.1f} us  ({mb:.0f} MB -> {mb / t * 1e-3 * 1e3:.2f} TB/s)".replace("TB/s", "GB/s x1e3"))
ref = (dy.double() @ w.double()) * (mask > 0)
L.linear_bwd(x, K, w, dy, N, None, None, dx, K, mask, K, M, K, N, 1, stream_ptr()); torch.cuda.synchronize()
print("rel err", float((dx.double() - ref).norm() / ref.norm()))
