"""Time the fused tcgen05 PointNet forward with individual epilogue passes knocked out (attribution)."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from bench import WORKLOADS, build_engine, flops_per_point
from pointcloud_rl_b200._lib import lib, stream_ptr
from pointcloud_rl_b200.synthetic import synthetic_batch

w = WORKLOADS["drq_maniskill_pn_jitter"]
eng, spec = build_engine(w, "bf16", "cuda:0", 0)
eng.upload_batch(synthetic_batch(0, w["B"], w["N"], w["A"], n_seg=w["n_seg"], state_dim=w["S"]))
L = eng.L
st = stream_ptr()
eng._pack_weights(st)
eng._stage("next_obs", "next", eng.k, 1, None, 1, st)
c1, c2, c3 = spec.widths
R = eng.R
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
def run(flags, iters=10):
    L.cdll.pcrl_debug_set_fwd_flags(ctypes.c_int(flags))
    ts = []
    for i in range(3 + iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        L.pointnet_fwd_bf16(eng.w["xh_next"], R, spec.n_points, spec.NP, eng.w["wpack"], c1, c2, c3, spec.ln_eps,
                            eng.w["pool_keys_next"], eng.w["pooled_next"], None, st)
        e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.mean(ts[3:]))
tiles = R * spec.NP // 128
for flags, name in [(0, "full"), (0, "full"), (8, "EXPERIMENT: no gamma/beta in L2 pass-2 (wrong results)"), (1, "no L2 pass-2"), 
                    (5, "no L2 epilogue at all"), (5 + 32 + 64, "sync skeleton + MMA")]:
    ms = run(flags)
    print(f"flags {flags:2d} {name:55s} {ms*1e3:8.1f} us   {ms*1e-3*1.9e9/ (tiles/148):8.0f} cyc/tile")
L.cdll.pcrl_debug_set_fwd_flags(ctypes.c_int(0))

# ---- round-trip trace of CTA 0 (clock64): skeleton mode and full mode
for flags in ((101 + 128, 128) if os.environ.get('PCRL_TRACE') else ()):
    L.cdll.pcrl_debug_set_fwd_flags(ctypes.c_int(flags))
    L.pointnet_fwd_bf16(eng.w["xh_next"], R, spec.n_points, spec.NP, eng.w["wpack"], c1, c2, c3, spec.ln_eps,
                        eng.w["pool_keys_next"], eng.w["pooled_next"], None, st)
    torch.cuda.synchronize()
    buf = (ctypes.c_longlong * 4096)()
    n = L.cdll.pcrl_debug_get_trace(buf, 2048)
    ev = sorted([(buf[2 * i + 1], buf[2 * i]) for i in range(n) if buf[2 * i + 1] != 0])
    t0 = ev[0][0]
    print(f"--- trace flags={flags}: {n} events; (cycles since first, event) 1xx=MMA ready-detected 2xx=MMA committed 3xx=epilogue woke 4xx=epilogue arrived; tens digit = slot, units = layer")
    prev = {}
    for t, e in ev[60:150]:
        print(f"{t - t0:8d} {e}")
L.cdll.pcrl_debug_set_fwd_flags(ctypes.c_int(0))
