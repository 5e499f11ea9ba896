import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pointcloud_rl_b200._lib import lib, stream_ptr
L = lib()
def rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).norm() / (b.norm() + 1e-12))
g = torch.Generator().manual_seed(3)
for (M,K,N) in [(256,1024,1),(256,1024,2),(256,1024,16),(256,1024,17),(256,64,1)]:
    x, w, b = torch.randn(M, K, generator=g), torch.randn(N, K, generator=g) / K**0.5, torch.randn(N, generator=g)
    y = torch.empty(M, N, device="cuda")
    L.linear_fwd(x.cuda(), K, w.cuda(), b.cuda(), y, N, M, K, N, 0, 1, stream_ptr())
    ref = x.double() @ w.double().t() + b.double()
    print(M,K,N, "fwd rel", rel(y, ref), y[:3,0].tolist(), ref[:3,0].tolist())
