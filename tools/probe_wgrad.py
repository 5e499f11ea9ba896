"""Weight-gradient-shaped GEMM (contraction over ~100 k rows): operand-major variants and split-K factors, timed with L2
flushed.  C[256,128] = A^T B with A [K,256], B [K,128] row-major (MN-major operands) vs pre-transposed K-major copies."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pointcloud_rl_b200._lib import lib, stream_ptr

L = lib()
dev = "cuda"
K, M, N = 100000, 256, 128
A = torch.randn(K, M, device=dev); B = torch.randn(K, N, device=dev)
At = A.t().contiguous(); Bt = B.t().contiguous()     # [M,K], [N,K] K-major
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
ref = (A.double().t() @ B.double())
def run(a, lda, a_mn, b, ldb, b_mn, sk):
    C = torch.zeros(M, N, device=dev)
    ts = []
    for i in range(6):
        C.zero_(); flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); L.gemm_tf32(a, lda, a_mn, b, ldb, b_mn, None, C, N, M, N, K, 0, 2, sk, stream_ptr()); e1.record()
        torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1) * 1e3)
    err = float((C.double() - ref).norm() / ref.norm())
    return min(ts[2:]), err
for sk in (37, 74, 148, 296):
    t, e = run(A, M, 1, B, N, 1, sk); print(f"MN/MN split_k={sk:3d}: {t:7.1f} us  rel {e:.1e}")
    t, e = run(At, K, 0, Bt, K, 0, sk); print(f"K /K  split_k={sk:3d}: {t:7.1f} us  rel {e:.1e}")
    t, e = run(A, M, 1, Bt, K, 0, sk); print(f"MN/K  split_k={sk:3d}: {t:7.1f} us  rel {e:.1e}")
