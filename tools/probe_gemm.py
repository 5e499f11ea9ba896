"""Diagnostic sweep of the tcgen05 TF32 GEMM over operand majors and shapes (prints relative errors)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pointcloud_rl_b200._lib import lib, stream_ptr

L = lib()
def rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).norm() / (b.norm() + 1e-12))
pad4 = lambda n: (n + 3) // 4 * 4
for (M, N, K) in [(128, 32, 32), (128, 32, 8), (128, 128, 64), (512, 1024, 1024), (200, 72, 136)]:
    for a_mn in (0, 1):
        for b_mn in (0, 1):
            g = torch.Generator().manual_seed(1)
            A = torch.randn(M, K, generator=g); Bm = torch.randn(K, N, generator=g)
            ref = A.double() @ Bm.double()
            if a_mn:
                lda = pad4(M); Ad = torch.zeros(K, lda).cuda(); Ad[:, :M] = A.t().cuda()
            else:
                lda = pad4(K); Ad = torch.zeros(M, lda).cuda(); Ad[:, :K] = A.cuda()
            if b_mn:
                ldb = pad4(N); Bd = torch.zeros(K, ldb).cuda(); Bd[:, :N] = Bm.cuda()
            else:
                ldb = pad4(K); Bd = torch.zeros(N, ldb).cuda(); Bd[:, :K] = Bm.t().cuda()
            C = torch.zeros(M, N, device="cuda")
            L.gemm_tf32(Ad, lda, a_mn, Bd, ldb, b_mn, None, C, N, M, N, K, 0, 0, 1, stream_ptr())
            torch.cuda.synchronize()
            print(f"M{M} N{N} K{K} a_mn{a_mn} b_mn{b_mn}: rel {rel(C, ref):.4f}  |C| {float(C.abs().max()):.3f} |ref| {float(ref.abs().max()):.3f}")
