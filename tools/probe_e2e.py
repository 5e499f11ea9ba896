"""Where the end-to-end time of agent.update_parameters(memory, updates) goes (host numpy batches): per-phase wall clock."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from bench import WORKLOADS, RotatingMemory, build_bench_agent
from pointcloud_rl_b200.data import unwrap
from pointcloud_rl_b200.synthetic import synthetic_batch

w = WORKLOADS["drq_maniskill_pn_jitter"]
agent = build_bench_agent(w, "bf16", "cuda:0", 0)
batches = [synthetic_batch(i, w["B"], w["N"], w["A"], n_seg=w["n_seg"], state_dim=w["S"]) for i in range(4)]
mem = RotatingMemory(batches)
for u in range(1, 9):
    agent.update_parameters(mem, u)
eng = agent.engine
torch.cuda.synchronize()

def timeit(fn, n=50):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for i in range(n): fn(i)
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / n * 1e3

print("update_parameters, result read at once   ms:", timeit(lambda i: dict(agent.update_parameters(mem, i + 1))))
prev = [None]
def late(i):
    cur = agent.update_parameters(mem, i + 1)
    if prev[0] is not None:
        prev[0]["drq/critic_loss"]
    prev[0] = cur
print("update_parameters, result read a step late ms:", timeit(late))
agent.engine.flush_scalars()
print("memory.sample + unwrap            ms:", timeit(lambda i: unwrap(mem.sample(256))))
b = unwrap(mem.sample(256))
print("upload_batch (stage + H2D enqueue) ms:", timeit(lambda i: eng.upload_batch(b)))
def host_only(i):
    for key, _s, _d, off, nb in eng._batch_layout:
        src = eng._host_leaf(b, key); dst = eng._pinned_np[0][key]; np.copyto(dst, src.reshape(dst.shape), casting="unsafe")
print("  numpy copies only, 1 thread     ms:", timeit(host_only))
print("  H2D of the pinned buffer only   ms:", timeit(lambda i: eng.raw_flat.copy_(eng._pinned[0], non_blocking=True)))
print("update_graphed only               ms:", timeit(lambda i: eng.update_graphed(i + 1)))
print("update_graphed + read_scalars     ms:", timeit(lambda i: (eng.update_graphed(i + 1), eng.read_scalars(i + 1))))
print("upload + update + scalars         ms:", timeit(lambda i: (eng.upload_batch(b), eng.update_graphed(i + 1), eng.read_scalars(i + 1))))
print("staging memcpy threads:", eng._copy_threads, " cpu_count:", os.cpu_count(), " torch threads:", torch.get_num_threads())
