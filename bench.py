#!/usr/bin/env python
"""Benchmark of the PointNet SAC/DrQ update path (BASELINE.json metric) on N B200s of one node.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the UNMODIFIED reference on the host CPU cores

A "step" is one `agent.update_parameters(memory, updates)` on one synthetic replay batch.  Default workload:
the DrQ ManiSkill `pn_jitter` shape (BASELINE config 2: B=256 x N=1200 x (xyz+rgb+1 seg mask) + 106-d agent
state, 22-d actions, num_aug=2, jitter +-0.01); even `updates` also run the actor/alpha/Polyak branches.
Other workloads: `sac_dmc_pn` (config 1), `encoder_fwd` (config 3, forward only), `sac_wide` (config 5).
Multi-GPU is data-parallel weak scaling (every rank its own batch, gradient all-reduce over NCCL), the
reference's own DDP semantics (run_rl.py:292-295).

Prints ONE JSON line (rank 0):
  value        distinct encoded points/s over all ranks, batches resident in HBM, device-timed (CUDA events);
  e2e          the same through the PUBLIC call -- an agent built by `build_agent` from the packaged config file,
               `agent.update_parameters(memory, updates)` with a host `memory` handing out fresh numpy batches:
               flatten + pinned staging + H2D + update + scalar D2H all inside the timed region; `e2e.device_ring`
               is the same call with the device-resident replay ring (only the sampled indices cross PCIe);
  roofline     the dominant kernel (fused PointNet forward) against the measured tensor peak;
  cpu_baseline the unmodified reference's update_parameters on this box's host cores (bounded sample);
  torch_cuda   the unmodified reference on this GPU through torch eager (.to("cuda")) -- the same-box comparator.
"""
import argparse
import contextlib
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # BASELINE.json configs[1]: DrQ configs/mfrl/drq/maniskill/pn_jitter.py, MoveBucket-shaped batch
    "drq_maniskill_pn_jitter": dict(algo="drq", cfg="mfrl/drq/maniskill/pn_jitter.py", B=256, N=1200, n_seg=1, n_pos=0,
                                    S=106, A=22, widths=(128, 128, 256), D=128, hidden=(1024, 1024), num_aug=2,
                                    aug="jitter", aug_lo=-0.01, aug_hi=0.01, gamma=0.95, zero_out_logstd=True),
    # BASELINE.json configs[0]: SAC configs/mfrl/sac/dm_control/pn.py
    "sac_dmc_pn": dict(algo="sac", cfg="mfrl/sac/dm_control/pn.py", B=128, N=1024, n_seg=0, n_pos=0, S=0, A=6,
                       widths=(64, 128, 256), D=50, hidden=(1024, 1024), num_aug=1, aug=None, aug_lo=0.0, aug_hi=0.0,
                       gamma=0.99, zero_out_logstd=False),
    # BASELINE.json configs[4]: wide PointNet (pointnet.py:81 default mlp_spec) SAC update, N=16384, B=512
    "sac_wide": dict(algo="sac", cfg="mfrl/sac/dm_control/pn.py", B=512, N=16384, n_seg=0, n_pos=0, S=0, A=6,
                     widths=(64, 128, 1024), D=256, hidden=(1024, 1024), num_aug=1, aug=None, aug_lo=0.0, aug_hi=0.0,
                     gamma=0.99, zero_out_logstd=False, cpu_sample_B=2,  # the reference's activations alone are 34 GB at B=512
                     overrides={"actor_cfg.nn_cfg.visual_nn_cfg.mlp_spec": [64, 128, 1024],
                                "actor_cfg.nn_cfg.visual_nn_cfg.out_channels": 256,
                                "actor_cfg.nn_cfg.mlp_cfg.mlp_spec": ["256", 1024, 1024, "action_shape * 2"],
                                "critic_cfg.nn_cfg.mlp_cfg.mlp_spec": ["256 + action_shape", 1024, 1024, 1]}),
    # BASELINE.json configs[2]: PointNet encoder forward only (rollout / actor path), ManiSkill PointNet
    "encoder_fwd": dict(algo="encode", B=4096, N=1200, n_seg=1, n_pos=0, S=106, A=22, widths=(128, 128, 256), D=128,
                        hidden=(1024, 1024), num_aug=1, aug=None),
}


def flops_per_point(C, widths):
    c1, c2, c3 = widths
    return 2 * (C * c1 + c1 * c2 + c2 * c3)


def encoded_points_per_update(w):
    if w["algo"] == "encode":
        return w["B"] * w["N"]
    k = w["num_aug"] if w["algo"] == "drq" else 1
    return (2 * k + 1) * w["B"] * w["N"]  # next_obs (kB) + obs (kB) + actor obs (B): SURVEY.md section 8d


def config_of(args, w, world):
    """The `config` object: identical in the native and the reference arm (same workload, same shapes)."""
    C = 6 + w["n_seg"] + w["n_pos"]
    return {"workload": args.workload, "batch_per_gpu": w["B"], "points": w["N"], "channels": C,
            "widths": list(w["widths"]), "num_aug": w["num_aug"] if w["algo"] == "drq" else 1,
            "parallelism": f"dp{world}",
            "l2": "4 distinct batches rotated; the per-step working set (staged points, activations of the compacted "
                  "backward, 27 MB of weights + Adam state) exceeds the 126 MB L2; the kernel-alone roofline timing "
                  "flushes L2 (256 MB memset) between launches"}


def load_peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        return {}


# ------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    """SM clock + throttle reasons of one GPU, sampled DURING the timed region by a background thread through NVML
    (in-process: a polling `nvidia-smi` child per rank takes driver locks often enough to slow the ranks it watches).
    Falls back to `nvidia-smi -lms` when the NVML binding is missing."""
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
    PERIOD_S = 0.1

    def __init__(self, gpu_index):
        self.gpu, self.rows, self.proc = gpu_index, [], None
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thread = None
        self._nvml = None

    def _physical_index(self):
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        if vis:
            try:
                return int(vis.split(",")[self.gpu])
            except Exception:
                pass
        return self.gpu

    def start(self):
        try:
            import pynvml

            pynvml.nvmlInit()
            self._nvml = pynvml
            self._handle = pynvml.nvmlDeviceGetHandleByIndex(self._physical_index())
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self._handle, pynvml.NVML_CLOCK_SM))
            self._thread = threading.Thread(target=self._poll_nvml, daemon=True)
            self._thread.start()
            return
        except Exception:
            self._nvml = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.gpu}", f"--query-gpu={self.QUERY}",
                                          "--format=csv,noheader,nounits", "-lms", "200"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _poll_nvml(self):
        nv = self._nvml
        flags = {"hw_slowdown": getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8),
                 "hw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
                 "sw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20),
                 "sw_power_cap": getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4)}
        get_reasons = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or nv.nvmlDeviceGetCurrentClocksThrottleReasons
        while not self._stop.is_set():
            try:
                self.samples.append(float(nv.nvmlDeviceGetClockInfo(self._handle, nv.NVML_CLOCK_SM)))
                mask = int(get_reasons(self._handle))
                for name, bit in flags.items():
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            self._stop.wait(self.PERIOD_S)

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self._nvml is not None:
            self._stop.set()
            self._thread.join(timeout=5)
            return {"sm_mhz": float(np.median(self.samples)) if self.samples else None, "sm_max_mhz": self.max_mhz,
                    "reasons": sorted(self.reasons), "samples": len(self.samples), "source": "nvml"}
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:  # nvidia-smi holds driver locks while it polls: make sure it is gone before anything else is timed
            self.proc.wait(timeout=10)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                mx = float(r[2])
                for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if val.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm), "source": "nvidia-smi"}


# ------------------------------------------------------------------------------------------ reference legs
class _Quiet(contextlib.AbstractContextManager):
    """The reference prints import-time notices to stdout; bench's stdout carries exactly one JSON line."""

    def __enter__(self):
        self._cm = contextlib.redirect_stdout(sys.stderr)
        self._cm.__enter__()
        return self

    def __exit__(self, *exc):
        return self._cm.__exit__(*exc)


def reference_seconds_per_update(name, w, device, steps, warmup, threads=None, batch_size=None):
    """The unmodified reference (oracle/_ref, staged by oracle/build_ref.py) when present -> kind "reference";
    else the oracle port -> kind "port" (CPU only)."""
    from oracle.ref_loader import reference_available

    if reference_available() and name in ("drq_maniskill_pn_jitter", "sac_dmc_pn"):
        from oracle.ref_bench import time_reference_updates

        with _Quiet():
            t, ret = time_reference_updates(name, w, device, steps, warmup, threads=threads, batch_size=batch_size)
        return t, "reference", ret
    if torch.device(device).type != "cpu":
        raise RuntimeError("reference not staged: no torch-CUDA comparator")
    return run_port_update(w, batch_size or w["B"], steps, warmup, threads or os.cpu_count() or 1), "port", None


def run_port_update(w, B_sample, n_steps, n_warm, threads):
    """Fallback CPU arm: the oracle port (pinned to the reference's golden vectors)."""
    from oracle import pointnet_sac_oracle as O

    torch.set_num_threads(threads)
    C = 6 + w["n_seg"] + w["n_pos"]
    params = O.init_params(0, C, w["widths"], w["D"], w["S"], w["A"], hidden=w["hidden"][0],
                           zero_out_logstd=w["zero_out_logstd"])
    state = O.new_state(params)
    batch = O.synthetic_batch(0, B_sample, w["N"], w["A"], n_seg=w["n_seg"], n_pos=w["n_pos"], state_dim=w["S"])
    k = w["num_aug"] if w["algo"] == "drq" else 1
    hp = dict(algo=w["algo"], gamma=w["gamma"], num_aug=k, aug=w["aug"])
    g = torch.Generator().manual_seed(0)

    def noise():
        n = {"eps_next": torch.randn(B_sample * k, w["A"], generator=g), "eps_pi": torch.randn(B_sample, w["A"], generator=g)}
        if w["aug"] == "jitter":
            for which in ("obs", "next"):
                n[f"jitter_{which}"] = w["aug_lo"] + (w["aug_hi"] - w["aug_lo"]) * torch.rand(B_sample * k, 3, w["N"], generator=g)
        return n

    times = []
    for u in range(1, n_warm + n_steps + 1):
        nz = noise()
        t0 = time.perf_counter()
        O.update(state, batch, u, hp, nz)
        if u > n_warm:
            times.append(time.perf_counter() - t0)
    return float(np.mean(times))


def reference_arm(args, w, rank, world):
    """`--impl reference`: the reference's own CPU implementation of the step, all host threads, on the native arm's
    config.  Full batch whenever K steps of it fit in ~4 minutes (K <= ~30 on a 16-core box); else the largest
    power-of-two slice of the batch that does, scaled by B / B_slice (per-sample work dominates) and said so."""
    if rank != 0:
        return
    if w["algo"] == "encode":
        print(json.dumps({"impl": "reference", "unavailable": "encoder_fwd has no reference arm; see tools/encoder_sweep.py"}))
        return
    cores = os.cpu_count() or 1
    # one probe update at 1/8 of the batch sizes the slice (cost is linear in B to within a few percent)
    B = w["B"]
    B_cap = int(w.get("cpu_sample_B", B))
    B_probe = min(B_cap, max(8, B // 8))
    t_probe, kind, _ = reference_seconds_per_update(args.workload, w, "cpu", 1, 1, threads=cores, batch_size=B_probe)
    t_full_est = t_probe * B / B_probe
    B_s = B_cap
    while B_s > min(8, B_cap) and (args.steps + args.warmup) * t_full_est * B_s / B > 270.0:
        B_s //= 2
    t, kind, ret = reference_seconds_per_update(args.workload, w, "cpu", args.steps, args.warmup, threads=cores, batch_size=B_s)
    t_full = t * (B / B_s)
    pts = encoded_points_per_update(w)
    value = pts / t_full
    what = "the unmodified reference (pyrl DrQ/SAC.update_parameters from oracle/_ref)" if kind == "reference" else "oracle port"
    sample = (f"{what}, {args.steps} timed updates of the full B={B} batch" if B_s == B else
              f"{what}, {args.steps} timed updates of a B={B_s} slice (1/{B // B_s} of the batch) scaled x{B // B_s}")
    line = {
        "impl": "reference", "metric": "update_encoded_points_per_s", "value": value, "unit": "points/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": t_full * 1e3,
        "steps_per_s": 1.0 / t_full, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "config": config_of(args, w, world),
        "cpu_baseline": {"value": value, "unit": "points/s", "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": "points/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "timed_region_s": t * args.steps,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------ GPU arm
class RotatingMemory:
    """Host `memory`: `sample(n)` hands out fresh numpy batches in turn (what ReplayMemory.sample returns,
    replay_buffer.py:297-322) wrapped in a DictArray; nothing is pinned or uploaded ahead of the call."""

    def __init__(self, batches):
        from pointcloud_rl_b200.data import DictArray

        self._wrap, self.batches, self.i = DictArray, batches, 0

    def sample(self, n):
        b = self.batches[self.i % len(self.batches)]
        self.i += 1
        return self._wrap(b)


def build_engine(w, precision, device, seed):
    """Bare UpdateEngine of a workload (tools/*.py: kernel probes and timelines below the agent API)."""
    from pointcloud_rl_b200.engine import HyperParams, PathSpec, UpdateEngine
    from pointcloud_rl_b200.synthetic import init_params

    spec = PathSpec(n_points=w["N"], action_dim=w["A"], state_dim=w["S"], n_pos=w["n_pos"], n_seg=w["n_seg"],
                    widths=w["widths"], out_dim=w["D"], hidden=w["hidden"])
    k = w["num_aug"] if w["algo"] == "drq" else 1
    hp = HyperParams(algo=w["algo"], gamma=w["gamma"], num_aug=k, aug=w["aug"], aug_lo=w["aug_lo"], aug_hi=w["aug_hi"])
    eng = UpdateEngine(spec, hp, batch_size=w["B"], device=device, precision=precision, seed=seed)
    eng.load_params(init_params(0, spec, zero_out_logstd=w["zero_out_logstd"]))
    eng.prime_alpha()
    return eng, spec


def build_bench_agent(w, precision, device, seed, use_graph=True):
    from pointcloud_rl_b200.synthetic import make_agent

    obs_shape = {"xyz": [3, w["N"]], "rgb": [3, w["N"]]}
    if w["n_pos"]:
        obs_shape["pos_encoding"] = [w["n_pos"], w["N"]]
    if w["n_seg"]:
        obs_shape["seg"] = [w["n_seg"], w["N"]]
    if w["S"]:
        obs_shape["agent"] = w["S"]
    torch.manual_seed(0)  # same initial weights on every rank (to_ddp broadcasts rank 0's anyway)
    agent = make_agent(w["cfg"], obs_shape, w["A"], overrides=w.get("overrides"), batch_size=w["B"], precision=precision,
                       use_cuda_graph=use_graph, seed=seed)
    return agent.to(device)


def time_dominant_kernel(eng, spec, w, iters=20):
    """CUDA-event timing of the fused PointNet forward kernel alone (the dominant kernel), for the roofline."""
    L = eng.L
    from pointcloud_rl_b200._lib import stream_ptr

    R = eng.R
    c1, c2, c3 = spec.widths
    st = stream_ptr()
    eng._pack_weights(st)
    eng._stage("next_obs", "next", eng.k, 1 if w["aug"] == "jitter" else 0, None, 1, st)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=eng.device)
    evs = []
    for _ in range(3 + iters):
        flush.zero_()  # > L2 (126 MB): the point tiles come from HBM every launch
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        eng._encode_points("next", R, False, st)
        e1.record()
        evs.append((e0, e1))
    torch.cuda.synchronize()
    ms = float(np.mean([a.elapsed_time(b) for a, b in evs[3:]]))
    flops = float(R) * spec.n_points * flops_per_point(spec.C, spec.widths)  # algorithmic: real points only
    return ms, flops


def measure_tf32_peak(device):
    """TF32 tensor peak with MEASURED_PEAKS.json's own method (torch.matmul 8192^3, best of 10 = burst; back to back
    for ~2 s = sustained): the denominator for the TF32 tcgen05 GEMMs of the MLP heads and the backward."""
    n = 8192
    old = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = True
    try:
        a = torch.randn(n, n, device=device)
        b = torch.randn(n, n, device=device)
        for _ in range(3):
            a @ b
        best = 1e9
        for _ in range(10):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            a @ b
            e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = max(10, int(2000.0 / best))
        e0.record()
        for _ in range(reps):
            a @ b
        e1.record()
        torch.cuda.synchronize()
        sustained = e0.elapsed_time(e1) / reps
        fl = 2.0 * n ** 3
        return {"tf32_tflops": fl / best / 1e9, "tf32_tflops_sustained": fl / sustained / 1e9,
                "how": "torch.matmul fp32 with allow_tf32 (cuBLAS), 8192^3: best of 10 (burst), back to back ~2 s (sustained)"}
    finally:
        torch.backends.cuda.matmul.allow_tf32 = old


def native_arm(args, w, rank, world, local_rank):
    import torch.distributed as dist

    from pointcloud_rl_b200.synthetic import synthetic_batch

    torch.cuda.set_device(local_rank)
    device = f"cuda:{local_rank}"
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(device))
    if w["algo"] == "encode":
        return encoder_arm(args, w, device)
    agent = build_bench_agent(w, args.dtype, device, seed=1234, use_graph=not args.no_graph)
    if world > 1:
        agent.to_ddp(device_ids=[local_rank])  # broadcasts rank 0's state, per-rank Philox stream, NCCL all-reduce

    # distinct batches: pageable numpy for the public call, resident in HBM for `value`
    n_pool = 4
    host_batches = [synthetic_batch(100 * rank + i, w["B"], w["N"], w["A"], n_seg=w["n_seg"], n_pos=w["n_pos"],
                                    state_dim=w["S"]) for i in range(n_pool)]
    memory = RotatingMemory(host_batches)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- warm-up through the public call (builds the engine, captures the CUDA graphs)
    for u in range(1, max(args.warmup, 4) + 1):
        ret = agent.update_parameters(memory, u)
    eng, spec = agent.engine, agent.engine.spec
    # from here on a tf32 GEMM whose operands are not TMA-addressable is an ERROR, not a silent FFMA fallback; the warm-up
    # above ran every shape of the workload once in the permissive mode, so the counter says whether any fell back
    fallbacks_warmup = int(eng.L.tf32_fallbacks())
    eng.L.set_strict_tf32(1 if fallbacks_warmup == 0 else 0)
    barrier()
    step_fn = eng.update_graphed if not args.no_graph else eng.update

    if args.profile_window:
        # `ncu --profile-from-start off ...`: exactly two updates (one critic-only, one with the actor/alpha/Polyak
        # branches) inside the cudaProfilerStart/Stop window, launched eagerly so every kernel is visible
        torch.cuda.profiler.start()
        eng.update(1)
        eng.update(2)
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
        _finish(world, [eng])
        return

    # ---- device-resident throughput (`value`)
    resident = [eng.to_device_batch(eng.make_pinned_batch(b)) for b in host_batches]
    eng.set_batch_device(resident[0])
    step_fn(1)
    sampler = ClockSampler(local_rank) if rank == 0 else None  # one watcher: rank 0's GPU
    if sampler:
        sampler.start()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches_before = eng.L.launches
    graph_kernels = 0
    e0.record()
    for i in range(args.steps):
        eng.set_batch_device(resident[i % n_pool])
        graph_kernels += step_fn(i + 1) or 0
    e1.record()
    barrier()
    dev_ms = e0.elapsed_time(e1)
    launches_eager = eng.L.launches - launches_before
    clocks = sampler.stop() if sampler else None

    # ---- end to end through the public call: host numpy batch -> pinned staging -> H2D -> update -> scalars D2H.
    # update_parameters returns its scalars as a dict that waits for the device->host copy on first access; the loop
    # below is the usual logging pattern -- launch update i, then read the scalars of update i-1 -- so all K results are
    # read inside the timed region while the host stages the next batch under the running update.
    def run_public(mem):
        log, prev = [], None
        for i in range(args.steps):
            cur = agent.update_parameters(mem, i + 1)
            if prev is not None:
                log.append(prev[loss_key])
            prev = cur
        log.append(prev[loss_key])
        return prev, log

    loss_key = f"{w['algo']}/critic_loss"
    barrier()
    t0 = time.perf_counter()
    e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e2.record()
    ret, e2e_log = run_public(memory)
    e3.record()
    barrier()
    e2e_ms = e2.elapsed_time(e3)
    e2e_wall = (time.perf_counter() - t0) * 1e3
    assert len(e2e_log) == args.steps and all(np.isfinite(e2e_log))
    h2d_bytes = sum(n for *_, n in eng._batch_layout)
    d2h_bytes = eng.scalars.numel() * 4

    # ---- the same call on the device-resident replay ring (replay.py): only the index vector crosses PCIe
    from pointcloud_rl_b200.replay import DeviceReplayMemory

    ring = DeviceReplayMemory(n_pool * w["B"], device=device, seed=rank)
    for b in host_batches:
        ring.push_batch(b)
    dict(agent.update_parameters(ring, 1))
    barrier()
    e4, e5 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e4.record()
    run_public(ring)
    e5.record()
    barrier()
    ring_ms = e4.elapsed_time(e5)

    times = torch.tensor([dev_ms, e2e_ms, ring_ms], device=device, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
    dev_ms, e2e_ms, ring_ms = [float(x) for x in times.tolist()]
    _finish(world, [eng], keep=(rank == 0))
    if rank != 0:
        return

    # ---- roofline of the dominant kernel (rank 0, kernel alone, L2 flushed between launches)
    k_ms, k_flops = time_dominant_kernel(eng, spec, w)
    peaks = load_peaks()
    tensor_dtype = "bf16" if args.dtype == "bf16" else ("tf32" if args.dtype == "tf32" else "fp32")
    tf32 = measure_tf32_peak(device) if world == 1 else None
    if tensor_dtype == "bf16":
        peak, peak_src = float(peaks.get("bf16_tflops", 1590.0)), ("MEASURED_PEAKS.json bf16_tflops (burst)" if peaks else "fallback")
    elif tensor_dtype == "tf32":
        peak, peak_src = (tf32["tf32_tflops"], "measured here, cuBLAS TF32 8192^3 burst") if tf32 else (
            float(peaks.get("bf16_tflops", 1590.0)) / 2, "half of the bf16 peak")
    else:
        peak, peak_src = 72.0, "nominal fp32 FFMA (148 SM x 128 FMA x 2 x 1.9 GHz)"
    achieved = k_flops / (k_ms * 1e-3) / 1e12
    traffic = None
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "dominant_kernel_traffic.json"))).get(args.dtype)
    except Exception:
        pass
    roofline = {"bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                "traffic": traffic, "kernel": eng.dominant_kernel_name(), "kernel_ms": k_ms, "peak_source": peak_src}

    # ---- baselines on the same box (N = 1 only): the unmodified reference on the host cores and on this GPU
    cpu = cuda_ref = None
    pts1 = encoded_points_per_update(w)
    if world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        B_s = int(w.get("cpu_sample_B", w["B"]))
        try:
            t, kind, _ = reference_seconds_per_update(args.workload, w, "cpu", 2, 1, threads=cores, batch_size=B_s)
            t *= w["B"] / B_s
            what = "the unmodified reference (oracle/_ref)" if kind == "reference" else "oracle port"
            size = f"the full B={w['B']} batch" if B_s == w["B"] else f"a B={B_s} slice of the batch, scaled x{w['B'] // B_s}"
            cpu = {"value": pts1 / t, "unit": "points/s", "cores": cores, "kind": kind, "steps_per_s": 1.0 / t,
                   "sample": f"{what}, 1 warm-up + 2 timed updates of {size}"}
        except Exception as e:  # noqa: BLE001
            cpu = {"value": None, "unit": "points/s", "cores": cores, "kind": "unavailable", "sample": repr(e)[:200]}
        try:
            if B_s != w["B"]:
                raise RuntimeError("the reference's dense autograd does not fit this configuration on one GPU "
                                   "(34 GB per saved activation at B=512, N=16384, c3=1024)")
            t, kind, _ = reference_seconds_per_update(args.workload, w, device, 10, 3)
            cuda_ref = {"value": pts1 / t, "unit": "points/s", "ms_per_step": t * 1e3, "kind": kind,
                        "what": "the unmodified reference moved to this GPU with .to('cuda'), torch eager fp32 "
                                "(TF32 off: torch default), 3 warm-up + 10 timed update_parameters calls, wall clock"}
        except Exception as e:  # noqa: BLE001
            cuda_ref = {"value": None, "unavailable": repr(e)[:200]}

    pts = pts1 * world
    ms_step = dev_ms / args.steps
    line = {
        "metric": "update_encoded_points_per_s", "value": pts / (ms_step * 1e-3), "unit": "points/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step, "steps_per_s": 1e3 / ms_step,
        "higher_is_better": True, "scaling": "strong" if args.global_batch else "weak", "vs_baseline": None,
        "dtype": args.dtype, "data": "synthetic", "config": config_of(args, w, world), "cuda_graph": not args.no_graph,
        "clocks": clocks,
        "e2e": {"value": pts / (e2e_ms / args.steps * 1e-3), "unit": "points/s", "ms_per_step": e2e_ms / args.steps,
                "wall_ms_per_step": e2e_wall / args.steps, "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": d2h_bytes,
                "api": "build_agent(cfg.agent_cfg).to('cuda').update_parameters(memory, updates); memory.sample returns "
                       "pageable numpy batches; every update's scalars are read on the host one step later (while the next "
                       "update runs), all inside the timed region",
                "device_ring": {"value": pts / (ring_ms / args.steps * 1e-3), "ms_per_step": ring_ms / args.steps,
                                "h2d_bytes_per_step": 8 * w["B"], "api": "same call, DeviceReplayMemory"}},
        "gpu_launches": int(launches_eager) if args.no_graph else int(graph_kernels),
        "roofline": roofline, "cpu_baseline": cpu, "torch_cuda": cuda_ref, "tf32_peak": tf32,
        "tf32_fallbacks": int(eng.L.tf32_fallbacks()),
        "last_scalars": {k: round(float(v), 5) for k, v in ret.items()},
    }
    print(json.dumps(line), flush=True)


def encoder_arm(args, w, device):
    """BASELINE config 3: PointNet encoder forward only (stage + fused per-point MLP/max-pool + final Linear/LN) on
    B clouds already resident in HBM; a step = one encode; L2 flushed between steps."""
    from pointcloud_rl_b200.engine import PathSpec
    from pointcloud_rl_b200.networks import KernelRunner
    from pointcloud_rl_b200.synthetic import init_params, synthetic_obs

    B, N = args.clouds or w["B"], w["N"]
    spec = PathSpec(n_points=N, action_dim=w["A"], state_dim=w["S"], n_seg=w["n_seg"], widths=w["widths"], out_dim=w["D"])
    p = {k: v.to(device) for k, v in init_params(0, spec).items()}
    rs = np.random.RandomState(B)
    obs_host = synthetic_obs(rs, B, N, n_seg=w["n_seg"])
    obs_pin = {k: torch.from_numpy(np.ascontiguousarray(v.astype(np.uint8) if v.dtype == bool else v)).pin_memory()
               for k, v in obs_host.items()}
    obs = {k: v.to(device) for k, v in obs_pin.items()}
    run = KernelRunner(args.dtype)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=device)
    for _ in range(max(3, args.warmup)):
        run.encode(spec, p, obs)
    sampler = ClockSampler(0)
    sampler.start()
    torch.cuda.synchronize()
    n0 = run.L.launches
    ts = []
    for _ in range(args.steps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        run.encode(spec, p, obs)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    launches = run.L.launches - n0
    clocks = sampler.stop()
    ms = float(np.mean(ts))
    # end to end: pinned host observations -> H2D -> encode -> features back on the host
    out_host = torch.empty(B, w["D"]).pin_memory()
    t_e2e = []
    for _ in range(args.steps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        o = {k: v.to(device, non_blocking=True) for k, v in obs_pin.items()}
        out_host.copy_(run.encode(spec, p, o), non_blocking=True)
        e1.record()
        torch.cuda.synchronize()
        t_e2e.append(e0.elapsed_time(e1))
    ms_e2e = float(np.mean(t_e2e))
    pts = B * N
    F_pt = flops_per_point(spec.C, spec.widths)
    peaks = load_peaks()
    peak = float(peaks.get("bf16_tflops", 1590.0)) if args.dtype == "bf16" else 72.0
    achieved = pts * F_pt / (ms * 1e-3) / 1e12
    line = {"metric": "encoder_points_per_s", "value": pts / (ms * 1e-3), "unit": "points/s", "n_gpus": 1,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": args.dtype, "data": "synthetic",
            "config": {"workload": args.workload, "clouds": B, "points": N, "channels": spec.C, "widths": list(w["widths"]),
                       "l2": "flushed (256 MB memset) between steps"},
            "clocks": clocks,
            "e2e": {"value": pts / (ms_e2e * 1e-3), "unit": "points/s", "ms_per_step": ms_e2e,
                    "h2d_bytes_per_step": int(sum(v.numel() * v.element_size() for v in obs_pin.values())),
                    "d2h_bytes_per_step": B * w["D"] * 4},
            "gpu_launches": int(launches),
            "roofline": {"bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                         "traffic": None, "kernel": "whole encode (stage + fused forward + head)", "kernel_ms": ms},
            "cpu_baseline": None}
    print(json.dumps(line), flush=True)


def _finish(world, engines, keep=False):
    """Multi-rank teardown: the captured CUDA graphs hold NCCL kernels, so they are destroyed BEFORE the communicator
    (the other order blocks in ncclCommDestroy).  keep=True leaves the engine usable (eager launches) for rank 0's
    kernel-alone timing after the group is gone."""
    if world <= 1:
        return
    import torch.distributed as dist

    for eng in engines:
        eng.close()
        eng.allreduce = None
    torch.cuda.synchronize()
    dist.barrier()
    dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--dtype", default="bf16", choices=["bf16", "tf32", "fp32"])
    ap.add_argument("--workload", default="drq_maniskill_pn_jitter", choices=sorted(WORKLOADS))
    ap.add_argument("--clouds", type=int, default=0, help="encoder_fwd: number of clouds per encode (default 4096)")
    ap.add_argument("--batch", type=int, default=0, help="override the workload's per-GPU batch size")
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--profile-window", action="store_true", help="run 2 eager updates inside cudaProfilerStart/Stop and exit")
    ap.add_argument("--global-batch", type=int, default=0,
                    help="strong scaling (SURVEY.md section 8d row 4): split this global minibatch over the ranks "
                         "(e.g. 256 or 2048); default 0 = weak scaling, the workload's batch on every rank")
    ap.add_argument("--allreduce", default="nccl", choices=["nccl", "peer_memory"],
                    help="N > 1: gradient all-reduce through NCCL (default) or the NVLink peer-memory kernel "
                         "(pcrl_p2p_allreduce; experimental, DESIGN.md section 7)")
    args = ap.parse_args()
    # watchdog: a rank that is still here after PCRL_BENCH_WATCHDOG seconds (default 25 min; a default run needs ~2) dumps
    # every thread's Python stack and exits instead of hanging the box until somebody else's limit kills it
    watchdog = float(os.environ.get("PCRL_BENCH_WATCHDOG", "1500"))
    if watchdog > 0:
        import faulthandler

        faulthandler.dump_traceback_later(watchdog, exit=True)
    os.environ["PCRL_P2P_ALLREDUCE"] = "1" if args.allreduce == "peer_memory" else "0"
    w = dict(WORKLOADS[args.workload])
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    if args.batch:
        w["B"] = args.batch
    if args.global_batch:
        if args.global_batch % world:
            raise SystemExit("--global-batch must be divisible by the number of ranks")
        w["B"] = args.global_batch // world
    if args.impl == "reference":
        reference_arm(args, w, rank, world)
        return
    native_arm(args, w, rank, world, local_rank)


if __name__ == "__main__":
    main()
