"""Where does the end-to-end (host batch -> update -> scalars) time go?"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from bench import WORKLOADS, build_engine
from pointcloud_rl_b200.synthetic import synthetic_batch

w = WORKLOADS["drq_maniskill_pn_jitter"]
eng, spec = build_engine(w, "bf16", "cuda:0", 0)
batches = [synthetic_batch(i, w["B"], w["N"], w["A"], n_seg=w["n_seg"], state_dim=w["S"]) for i in range(4)]
pinned = [eng.make_pinned_batch(b) for b in batches]
print({k: (tuple(v.shape), v.dtype, v.is_pinned()) for k, v in pinned[0].items()})
cs = torch.cuda.Stream()
eng.upload_batch(batches[0])
for u in range(1, 5): eng.update_graphed(u)
torch.cuda.synchronize()
def timeit(fn, n=20):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for i in range(n): fn(i)
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / n * 1e3
# (a) H2D only
def h2d(i):
    ev, nb = eng.h2d_async(pinned[i % 4], i % 2, cs)
print("H2D only          ms:", timeit(h2d))
big = torch.empty(64 << 20, dtype=torch.uint8).pin_memory(); dbig = torch.empty_like(big, device="cuda")
print("H2D 64 MB pinned  GB/s:", 64 / 1024 / (timeit(lambda i: dbig.copy_(big, non_blocking=True), 10) / 1e3))
print("update only       ms:", timeit(lambda i: eng.update_graphed(i + 1)))
print("update+scalars    ms:", timeit(lambda i: (eng.update_graphed(i + 1), eng.read_scalars(i + 1))))
def adopt_only(i):
    ev, _ = eng.h2d_async(pinned[i % 4], i % 2, cs); eng.adopt(i % 2, ev)
print("h2d+adopt         ms:", timeit(adopt_only))
def full(i):
    global ev
    eng.adopt(i % 2, ev)
    cs.wait_stream(torch.cuda.current_stream())
    ev, _ = eng.h2d_async(pinned[(i + 1) % 4], (i + 1) % 2, cs)
    eng.update_graphed(i + 1)
    eng.read_scalars(i + 1)
ev, _ = eng.h2d_async(pinned[0], 0, cs)
print("full e2e loop     ms:", timeit(full))
print("full e2e loop     ms:", timeit(full))
def simple(i):
    eng.upload_batch(batches[i % 4]); eng.update_graphed(i + 1); eng.read_scalars(i + 1)
print("upload_batch(numpy)+update+scalars ms:", timeit(simple))
