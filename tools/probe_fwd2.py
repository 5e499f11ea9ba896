"""A/B of the two fused tcgen05 PointNet forwards (first generation: points on lanes + shared-memory transpose max;
second generation: transposed layer 2 + Gram-matrix variance): accuracy against the fp32 oracle on the golden
PointNet fixtures, then kernel time at BASELINE config 2 (512 clouds x 1200 points, L2 flushed).
Usage (GPU box): python tools/probe_fwd2.py"""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle import pointnet_sac_oracle as O
from pointcloud_rl_b200._lib import lib, stream_ptr
from tests.conftest import load_golden

L = lib()
st = stream_ptr()


def rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).norm() / b.norm())


def run_case(name, version, R_rep=1):
    g = load_golden(name)
    p = {k: torch.from_numpy(np.asarray(v)) for k, v in g["params"].items()}
    obs = {k: torch.from_numpy(np.asarray(v)) for k, v in g["obs"].items()}
    x = O.preprocess(obs)
    R, C, N = x.shape
    NP = (N + 127) // 128 * 128
    CP = 8 if C <= 8 else 16
    xf = torch.zeros(R, NP, CP, device="cuda")
    xh = torch.zeros(R * NP * 16, dtype=torch.bfloat16, device="cuda")
    seg, pos = obs.get("seg"), obs.get("pos_encoding")
    L.stage_points(obs["xyz"].cuda(), obs["rgb"].cuda(), 1, pos.cuda() if pos is not None else None,
                   0 if pos is None else pos.shape[1], seg.to(torch.uint8).cuda() if seg is not None else None,
                   0 if seg is None else seg.shape[1], R, N, 1, 0, 0.0, 0.0, None, 0, None, 0, xf, xh, CP, st)
    c1, c2, c3 = p["pn.w0"].shape[0], p["pn.w1"].shape[0], p["pn.w2"].shape[0]
    d = {k: v.cuda().contiguous() for k, v in p.items()}
    wpack = torch.zeros(int(L.pointnet_wpack_bytes(c1, c2, c3)), dtype=torch.uint8, device="cuda")
    L.pointnet_pack_weights(d["pn.w0"], d["pn.b0"], d["pn.w1"], d["pn.g1"], d["pn.be1"], d["pn.w2"], d["pn.g2"],
                            d["pn.be2"], C, c1, c2, c3, 1, wpack, st)
    keys = torch.zeros(R * c3, dtype=torch.int64, device="cuda")
    pooled = torch.empty(R, c3, device="cuda")
    argmax = torch.empty(R, c3, dtype=torch.int32, device="cuda")
    L.cdll.pcrl_debug_set_fwd_version(ctypes.c_int(version))
    L.pointnet_fwd_bf16(xh, R, N, NP, wpack, c1, c2, c3, 1e-6, keys, pooled, argmax, st)
    pooled2 = torch.empty_like(pooled)
    L.pointnet_fwd_bf16(xh, R, N, NP, wpack, c1, c2, c3, 1e-6, keys, pooled2, None, st)
    torch.cuda.synchronize()
    ref = torch.from_numpy(g["pooled"])
    h = O.pointnet_point_features(p, x)
    idx = argmax.cpu().long().clamp(0, N - 1)
    v_ours = torch.gather(h, 2, idx[..., None])[..., 0]
    mism = float((argmax.cpu().long() != torch.from_numpy(g["idx"])).float().mean())
    print(f"{name:24s} v{version}: pooled rel err {rel(pooled, ref):.3e}  values-only == argmax variant: {torch.equal(pooled, pooled2)}"
          f"  idx range [{int(argmax.min())},{int(argmax.max())}] of {N}  argmax mismatch {mism:.3f}"
          f"  worst (true max - true value at our argmax)/max: {float((ref - v_ours).max() / ref.max()):.3e}  keys zeroed: {int(keys.abs().max()) == 0}")


for name in ("pointnet_fwd_c7", "pointnet_fwd_c7_dup", "pointnet_fwd_c9_dmc"):
    for ver in (1, 2):
        run_case(name, ver)

# ---- timing at config 2
from bench import WORKLOADS, flops_per_point
from pointcloud_rl_b200.engine import HyperParams, PathSpec, UpdateEngine
from pointcloud_rl_b200.synthetic import init_params, synthetic_batch

w = WORKLOADS["drq_maniskill_pn_jitter"]
spec = PathSpec(n_points=w["N"], action_dim=w["A"], state_dim=w["S"], n_seg=w["n_seg"], widths=w["widths"], out_dim=w["D"])
eng = UpdateEngine(spec, HyperParams(algo="drq", num_aug=2, aug="jitter", aug_lo=-0.01, aug_hi=0.01), w["B"], precision="bf16")
eng.load_params(init_params(0, spec, zero_out_logstd=True))
eng.upload_batch(synthetic_batch(0, w["B"], w["N"], w["A"], n_seg=w["n_seg"], state_dim=w["S"]))
eng._pack_weights(st)
eng._stage("next_obs", "next", eng.k, 1, None, 1, st)
c1, c2, c3 = spec.widths
R = eng.R
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
tiles = R * spec.NP // 128
res = {}
for ver in (1, 2):
    L.cdll.pcrl_debug_set_fwd_version(ctypes.c_int(ver))
    for want in (False, True):
        ts = []
        for i in range(13):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            L.pointnet_fwd_bf16(eng.w["xh_next"], R, spec.n_points, spec.NP, eng.w["wpack"], c1, c2, c3, spec.ln_eps,
                                eng.w["pool_keys_next"], eng.w["pooled_next"], eng.w["argmax_obs"] if want else None, st)
            e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ms = float(np.mean(ts[3:]))
        fl = R * spec.n_points * flops_per_point(spec.C, spec.widths)
        res[(ver, want)] = eng.w["pooled_next"].clone()
        print(f"v{ver} argmax={int(want)}: {ms*1e3:7.1f} us  {ms*1e-3*1.965e9/(tiles/148):7.0f} cyc/tile  {fl/ms/1e9:7.1f} TFLOP/s  "
              f"= {fl/ms/1e9/1673*100:.1f} % of 1673")
L.cdll.pcrl_debug_set_fwd_version(ctypes.c_int(2))
for flags, name in [(1, "no pool math"), (2, "no Gram dot math"), (4, "no layer-1 normalise math"), (1 + 2 + 4, "no epilogue math at all"),
                    (1 + 16, "no pool math, no pool TMEM loads"), (2 + 4 + 8, "no front math, no front TMEM loads"),
                    (31, "sync skeleton + MMAs only")]:
    L.cdll.pcrl_debug_set_fwd2_flags(ctypes.c_int(flags))
    ts = []
    for i in range(8):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        L.pointnet_fwd_bf16(eng.w["xh_next"], R, spec.n_points, spec.NP, eng.w["wpack"], c1, c2, c3, spec.ln_eps,
                            eng.w["pool_keys_next"], eng.w["pooled_next"], None, st)
        e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ms = float(np.mean(ts[3:]))
    print(f"knock-out {flags:2d} {name:40s} {ms*1e3:7.1f} us  {ms*1e-3*1.965e9/(tiles/148):7.0f} cyc/tile")
# ---- clock64 trace of CTA 0 (needs a build with PCRL_NVCC_EXTRA=-DPCRL_FWD_TRACE)
if os.environ.get("PCRL_TRACE"):
    for flags in [int(x) for x in os.environ["PCRL_TRACE"].split(",")]:
        L.cdll.pcrl_debug_set_fwd2_flags(ctypes.c_int(flags | 256))
        L.pointnet_fwd_bf16(eng.w["xh_next"], R, spec.n_points, spec.NP, eng.w["wpack"], c1, c2, c3, spec.ln_eps,
                            eng.w["pool_keys_next"], eng.w["pooled_next"], None, st)
        torch.cuda.synchronize()
        buf = (ctypes.c_longlong * (6 * 1024))()
        L.cdll.pcrl_debug_get_trace2(buf)
        ev = []
        for reg in range(6):
            for i in range(512):
                e, t = buf[reg * 1024 + 2 * i], buf[reg * 1024 + 2 * i + 1]
                if t:
                    ev.append((t, reg, e))
        ev.sort()
        t0 = ev[0][0] if ev else 0
        names = ["iss0", "iss1", "issT", "frn0", "frn1", "pool"]
        print(f"--- trace, knock-out flags {flags}: role, tile, event (1xx issuer woke: 100 L0 / 110 L1 / 120 U, 220 U issued; 13b issT woke for block b;"
              " 300 F0, 301 HF, 310 E0 sent, 320 F1, 330 E1 sent, 340 FU, 350 EU sent; 400 pool got rstd, 41b block ready, 42b block done)")
        for t, reg, e in ev:
            tile = e // 1000
            if 12 <= tile <= 15:
                print(f"{t - t0:8d}  {names[reg]}  tile {tile:2d}  {e % 1000}")
L.cdll.pcrl_debug_set_fwd2_flags(ctypes.c_int(0))
print("config-2 pooled v2 vs v1 rel diff:", rel(res[(2, False)], res[(1, False)]))
L.cdll.pcrl_debug_set_fwd_version(ctypes.c_int(0))
