#!/usr/bin/env python
"""Benchmark of the PointNet SAC/DrQ update path (BASELINE.json metric) on N B200s of one node.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the reference algorithm on the host CPU cores

A "step" is one `update_parameters` on one synthetic replay batch of the DrQ ManiSkill `pn_jitter`
shape (BASELINE config 2: B=256 x N=1200 x (xyz+rgb+1 seg mask) + 106-d agent state, 22-d actions,
num_aug=2, jitter +-0.01); even `updates` also run the actor/alpha/Polyak branches (interval 2).
Multi-GPU is data-parallel weak scaling (every rank its own B=256 batch, gradient all-reduce over NCCL),
as the reference's own DDP semantics (run_rl.py:292-295).

Prints ONE JSON line (rank 0).  `value` = distinct encoded points/s over all ranks with the batches
resident in HBM; `e2e` = the same through the public agent API with HOST (pinned) batches, H2D copies
and the scalar read-back inside the timed region.
"""
import argparse
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
    "drq_maniskill_pn_jitter": dict(algo="drq", B=256, N=1200, n_seg=1, n_pos=0, S=106, A=22, widths=(128, 128, 256),
                                    D=128, hidden=(1024, 1024), num_aug=2, aug="jitter", aug_lo=-0.01, aug_hi=0.01,
                                    gamma=0.95, zero_out_logstd=True),
    # BASELINE.json configs[0]: SAC configs/mfrl/sac/dm_control/pn.py
    "sac_dmc_pn": dict(algo="sac", B=128, N=1024, n_seg=0, n_pos=0, S=0, A=6, widths=(64, 128, 256), D=50,
                       hidden=(1024, 1024), num_aug=1, aug=None, aug_lo=0.0, aug_hi=0.0, gamma=0.99,
                       zero_out_logstd=False),
}


def flops_per_point(C, widths):
    c1, c2, c3 = widths
    return 2 * (C * c1 + c1 * c2 + c2 * c3)


def encoded_points_per_update(w):
    k = w["num_aug"] if w["algo"] == "drq" else 1
    return (2 * k + 1) * w["B"] * w["N"]  # next_obs (kB) + obs (kB) + actor obs (B): SURVEY.md section 8d


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


# ------------------------------------------------------------------------------------------ CPU arm
def run_cpu_update(w, B_sample, n_steps, n_warm, threads):
    """The reference algorithm (oracle port, pinned to the reference's golden vectors) on the host cores."""
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
        dt = time.perf_counter() - t0
        if u > n_warm:
            times.append(dt)
    return float(np.mean(times))


def reference_arm(args, w, rank):
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    # bounded sample: a B/4 slice per step (as the native arm's cpu_baseline) unless K steps of it would run for more
    # than ~2 minutes (≈ 6 ms of host time per sample on this class of box), then a smaller power-of-two slice
    B_s = max(8, w["B"] // 4)
    while B_s > 8 and args.steps * 0.006 * B_s > 120.0:
        B_s //= 2
    t = run_cpu_update(w, B_s, args.steps, max(1, min(args.warmup, 2)), cores)
    t_full = t * (w["B"] / B_s)  # a full-batch step costs B/B_s sampled steps (per-sample work dominates)
    pts = encoded_points_per_update(w)
    value = pts / t_full
    sample = f"oracle port, {args.steps} timed updates of a B={B_s} slice (1/{w['B'] // B_s} of the batch) scaled x{w['B'] // B_s}"
    line = {
        "impl": "reference", "metric": "update_encoded_points_per_s", "value": value, "unit": "points/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": t_full * 1e3,
        "steps_per_s": 1.0 / t_full, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": args.workload, "batch_per_gpu": w["B"], "points": w["N"],
                   "channels": 6 + w["n_seg"] + w["n_pos"], "num_aug": w["num_aug"] if w["algo"] == "drq" else 1,
                   "parallelism": "host cores", "cuda_graph": False},
        "cpu_baseline": {"value": value, "unit": "points/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "points/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------ GPU arm
def build_engine(w, precision, device, seed):
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
        if eng.precision == "bf16":
            L.pointnet_fwd_bf16(eng.w["xh_next"], R, spec.n_points, spec.NP, eng.w["wpack"], c1, c2, c3, spec.ln_eps,
                                eng.w["pool_keys_next"], eng.w["pooled_next"], None, st)
        else:
            p = eng.p
            L.pointnet_fwd_f32(eng.w["xf_next"], R, spec.n_points, spec.NP, spec.CP, spec.C, p["pn.w0"], p["pn.b0"],
                               p["pn.w1"], p["pn.g1"], p["pn.be1"], p["pn.w2"], p["pn.g2"], p["pn.be2"], c1, c2, c3,
                               spec.ln_eps, eng.w["pooled_next"], None, eng.w["scratch"], eng.fwd_ws_bytes, st)
        e1.record()
        evs.append((e0, e1))
    torch.cuda.synchronize()
    ms = float(np.mean([a.elapsed_time(b) for a, b in evs[3:]]))
    flops = float(R) * spec.n_points * flops_per_point(spec.C, spec.widths)  # algorithmic: real points only
    return ms, flops


def native_arm(args, w, rank, world, local_rank):
    import torch.distributed as dist

    from pointcloud_rl_b200.synthetic import synthetic_batch

    torch.cuda.set_device(local_rank)
    device = f"cuda:{local_rank}"
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(device))
    eng, spec = build_engine(w, args.dtype, device, seed=1234 + rank)
    if world > 1:
        from pointcloud_rl_b200.dist import attach

        attach(eng, dist.group.WORLD)
    launches0 = eng.L.launches

    # a pool of distinct batches: resident in HBM for `value`, pinned on the host for `e2e`
    n_pool = 4
    host_batches = [synthetic_batch(100 * rank + i, w["B"], w["N"], w["A"], n_seg=w["n_seg"], n_pos=w["n_pos"],
                                    state_dim=w["S"]) for i in range(n_pool)]
    pinned = [eng.make_pinned_batch(b) for b in host_batches]
    resident = [eng.to_device_batch(pb) for pb in pinned]
    step_fn = eng.update_graphed if not args.no_graph else eng.update

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- warm-up (also captures the two CUDA graphs)
    eng.set_batch_device(resident[0])
    for u in range(1, max(args.warmup, 4) + 1):
        step_fn(u)
    barrier()

    if args.profile_window:
        # `ncu --profile-from-start off ...`: exactly two updates (one critic-only, one with the actor/alpha/Polyak
        # branches) inside the cudaProfilerStart/Stop window, launched eagerly so every kernel is visible
        torch.cuda.profiler.start()
        eng.update(1)
        eng.update(2)
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
        _finish(world)
        return

    # ---- device-resident throughput (`value`)
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

    # ---- end to end through the host-facing path: pinned batch -> H2D -> update -> scalar read-back
    copy_stream = torch.cuda.Stream(device=device)
    barrier()
    t0 = time.perf_counter()
    e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e2.record()
    ev, h2d_bytes = eng.h2d_async(pinned[0], 0, copy_stream)
    d2h_bytes = 0
    for i in range(args.steps):
        eng.adopt(i % 2, ev)
        step_fn(i + 1)
        if i + 1 < args.steps:
            # the other landing slot was last read by the adopt of step i-1, which has completed (per-step sync below)
            ev, _ = eng.h2d_async(pinned[(i + 1) % n_pool], (i + 1) % 2, copy_stream)  # overlaps update i
        ret = eng.read_scalars(i + 1, sync=True)  # the loss/metrics every update_parameters call returns
        d2h_bytes = eng.scalars.numel() * 4
    e3.record()
    barrier()
    e2e_ms = e2.elapsed_time(e3)
    e2e_wall = (time.perf_counter() - t0) * 1e3

    times = torch.tensor([dev_ms, e2e_ms], device=device, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
    dev_ms, e2e_ms = [float(x) for x in times.tolist()]

    if rank != 0:
        _finish(world)
        return

    # ---- roofline of the dominant kernel (rank 0, kernel alone, L2 flushed between launches)
    k_ms, k_flops = time_dominant_kernel(eng, spec, w)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    if args.dtype == "bf16":
        peak, peak_src = float(peaks.get("bf16_tflops", 1590.0)), ("measured" if peaks else "fallback")
    else:
        peak, peak_src = 72.0, "nominal fp32 FFMA (148 SM x 128 FMA x 2 x 1.9 GHz)"
    achieved = k_flops / (k_ms * 1e-3) / 1e12
    traffic = None
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "dominant_kernel_traffic.json"))).get(args.dtype)
    except Exception:
        pass
    roofline = {"bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                "traffic": traffic, "kernel": "pointnet_fwd_tc_kernel" if args.dtype == "bf16" else "pointnet_fwd_f32 chain",
                "kernel_ms": k_ms, "peak_source": peak_src}

    # ---- CPU baseline: the reference algorithm on this box's host cores, bounded sample
    cpu = None
    if not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        B_s = max(8, w["B"] // 4)
        n_cpu = 24  # ~10 s of host work on this box
        t = run_cpu_update(w, B_s, n_cpu, 1, cores) * (w["B"] / B_s)
        cpu = {"value": encoded_points_per_update(w) / t, "unit": "points/s", "cores": cores, "kind": "port",
               "sample": f"oracle port, {n_cpu} timed updates of a B={B_s} slice (1/{w['B'] // B_s} of the batch) scaled x{w['B'] // B_s}",
               "steps_per_s": 1.0 / t}

    pts = encoded_points_per_update(w) * world
    ms_step = dev_ms / args.steps
    line = {
        "metric": "update_encoded_points_per_s", "value": pts / (ms_step * 1e-3), "unit": "points/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step, "steps_per_s": 1e3 / ms_step,
        "higher_is_better": True, "scaling": "strong" if args.global_batch else "weak", "vs_baseline": None,
        "dtype": args.dtype, "data": "synthetic",
        "config": {"workload": args.workload, "batch_per_gpu": w["B"], "points": w["N"], "channels": spec.C,
                   "num_aug": eng.k, "parallelism": f"dp{world}", "cuda_graph": not args.no_graph,
                   "l2": "4 distinct resident batches rotated; per-step working set (staged points, activations of the compacted backward, 27 MB weights+Adam state) exceeds the 126 MB L2"},
        "clocks": clocks,
        "e2e": {"value": pts / (e2e_ms / args.steps * 1e-3), "unit": "points/s", "ms_per_step": e2e_ms / args.steps,
                "wall_ms_per_step": e2e_wall / args.steps, "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": d2h_bytes},
        "gpu_launches": int(launches_eager) if args.no_graph else int(graph_kernels),
        "roofline": roofline, "cpu_baseline": cpu, "last_scalars": {k: round(v, 5) for k, v in ret.items()},
    }
    print(json.dumps(line), flush=True)
    _finish(world)


def _finish(world):
    """Multi-rank exit: tearing the NCCL communicator down while captured CUDA graphs still reference its kernels
    can block forever, so ranks just flush and leave (exit code 0) once their own work is done."""
    if world > 1:
        sys.stdout.flush()
        sys.stderr.flush()
        torch.cuda.synchronize()
        os._exit(0)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=400)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--dtype", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--workload", default="drq_maniskill_pn_jitter", choices=sorted(WORKLOADS))
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--profile-window", action="store_true", help="run 2 eager updates inside cudaProfilerStart/Stop and exit")
    ap.add_argument("--global-batch", type=int, default=0,
                    help="strong scaling (SURVEY.md section 8d row 4): split this global minibatch over the ranks "
                         "(e.g. 256 or 2048); default 0 = weak scaling, the workload's batch on every rank")
    args = ap.parse_args()
    w = dict(WORKLOADS[args.workload])
    if args.global_batch:
        world_ = int(os.environ.get("WORLD_SIZE", 1))
        if args.global_batch % world_:
            raise SystemExit("--global-batch must be divisible by the number of ranks")
        w["B"] = args.global_batch // world_
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    if args.impl == "reference":
        reference_arm(args, w, rank)
        return
    native_arm(args, w, rank, world, local_rank)


if __name__ == "__main__":
    main()
