"""Times the UNMODIFIED reference's own `update_parameters` (pyrl/methods/mfrl/{sac,drq}.py) on a synthetic replay
batch -- on the host cores (bench.py --impl reference / cpu_baseline) or on the GPU through torch-CUDA eager (the
"same box" comparator SURVEY.md section 8d asks for).  TEST / BENCH INFRASTRUCTURE ONLY.

The agent is built exactly as run_rl.py builds it: Config.fromfile(configs/...) -> replace_placeholder_with_args ->
build_agent; `memory` is the minimal object update_parameters needs (`sample(n) -> DictArray`, sac.py:104).
"""
import os
import time

import numpy as np
import torch

from . import pointnet_sac_oracle as O
from .ref_loader import build_reference_agent, load_reference, reference_available

CONFIG_OF = {
    "drq_maniskill_pn_jitter": "configs/mfrl/drq/maniskill/pn_jitter.py",
    "sac_dmc_pn": "configs/mfrl/sac/dm_control/pn.py",
}


class RotatingMemory:
    """`memory.sample(n)`: hands out pre-generated synthetic batches in turn, as numpy (what ReplayMemory.sample
    returns, replay_buffer.py:297-322) wrapped in the reference's DictArray."""

    def __init__(self, ns, batches):
        self.ns, self.batches, self.i = ns, batches, 0

    def sample(self, n):
        b = self.batches[self.i % len(self.batches)]
        self.i += 1
        b = {k: (dict(v) if isinstance(v, dict) else v) for k, v in b.items()}
        b["prev_actions"] = np.zeros_like(b["actions"])
        b["episode_dones"] = b["dones"].copy()
        return self.ns.DictArray(b)


def obs_shape_of(obs):
    return {k: (list(v.shape[1:]) if v.ndim > 2 else int(v.shape[1])) for k, v in obs.items()}


def time_reference_updates(workload_name, w, device, steps, warmup, threads=None, batch_size=None, n_pool=2):
    """-> (seconds per update (mean over `steps` timed updates), last returned scalar dict).  Raises if the
    reference is not staged."""
    if not reference_available():
        raise RuntimeError("reference not staged (oracle/_ref)")
    if threads:
        torch.set_num_threads(threads)
    ns = load_reference()
    B = int(batch_size or w["B"])
    batches = [O.synthetic_batch(i, B, w["N"], w["A"], n_seg=w["n_seg"], n_pos=w["n_pos"], state_dim=w["S"])
               for i in range(n_pool)]
    torch.manual_seed(0)
    agent, _ = build_reference_agent(ns, CONFIG_OF[workload_name], obs_shape_of(batches[0]["obs"]), w["A"],
                                     {"batch_size": B})
    agent = agent.to(device)
    mem = RotatingMemory(ns, batches)
    cuda = torch.device(device).type == "cuda"
    ret = None
    for u in range(1, warmup + 1):
        ret = agent.update_parameters(mem, updates=u)
    if cuda:
        torch.cuda.synchronize()
    t0 = time.perf_counter()
    for u in range(warmup + 1, warmup + steps + 1):
        ret = agent.update_parameters(mem, updates=u)  # every call ends in .item() reads: synchronous on CUDA too
    if cuda:
        torch.cuda.synchronize()
    return (time.perf_counter() - t0) / steps, ret
