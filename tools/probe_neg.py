import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle import pointnet_sac_oracle as O
from tests.conftest import load_golden
from pointcloud_rl_b200._lib import lib, stream_ptr
L = lib(); sp = stream_ptr
g = load_golden("pointnet_fwd_c7")
p = {k: torch.from_numpy(np.asarray(v)) for k, v in g["params"].items()}
obs = {k: torch.from_numpy(np.asarray(v)) for k, v in g["obs"].items()}
for variant in ("base", "g1one", "g1all", "g1neg", "g2neg", "be1big"):
    q = dict(p)
    gen = torch.Generator().manual_seed(11)
    if variant == "g2neg":
        q["pn.g2"] = p["pn.g2"] * torch.where(torch.rand(p["pn.g2"].shape, generator=gen) < 0.4, -1.0, 1.0)
    if variant == "g1neg":
        q["pn.g1"] = p["pn.g1"] * torch.where(torch.rand(p["pn.g1"].shape, generator=gen) < 0.3, -1.0, 1.0)
    if variant == "g1one":
        q["pn.g1"] = p["pn.g1"].clone(); q["pn.g1"][3] *= -1
    if variant == "g1all":
        q["pn.g1"] = -p["pn.g1"]
    if variant == "be1big":
        q["pn.be1"] = p["pn.be1"] + 1.0
    x = O.preprocess(obs)
    _, ref, _ = O.pointnet_forward(q, x, return_pool=True)
    R, C, N = x.shape; NP = (N + 127) // 128 * 128
    xf = torch.zeros(R, NP, 8, device="cuda"); xh = torch.zeros(R * NP * 16, dtype=torch.bfloat16, device="cuda")
    L.stage_points(obs["xyz"].cuda(), obs["rgb"].cuda(), 1, None, 0, obs["seg"].to(torch.uint8).cuda(), 1, R, N, 1, 0, 0.0, 0.0, None, 0, None, 0, xf, xh, 8, sp())
    d = {k: v.cuda().contiguous() for k, v in q.items()}
    wpack = torch.zeros(int(L.pointnet_wpack_bytes(128, 128, 256)) + 4096, dtype=torch.uint8, device="cuda")
    L.pointnet_pack_weights(d["pn.w0"], d["pn.b0"], d["pn.w1"], d["pn.g1"], d["pn.be1"], d["pn.w2"], d["pn.g2"], d["pn.be2"], C, 128, 128, 256, 1, wpack, sp())
    keys = torch.zeros(R * 256, dtype=torch.int64, device="cuda"); pooled = torch.empty(R, 256, device="cuda")
    L.pointnet_fwd_bf16(xh, R, N, NP, wpack, 128, 128, 256, 1e-6, keys, pooled, None, sp())
    torch.cuda.synchronize()
    bf = lambda t: t.to(torch.bfloat16).float()
    hh = bf(torch.relu(O._conv1x1(bf(q["pn.w0"]), x, bf(q["pn.b0"]))))
    hh = bf(torch.relu(O._ln_channels(O._conv1x1(bf(q["pn.w1"]), hh), q["pn.g1"], q["pn.be1"], 1e-6)))
    emu = torch.relu(O._ln_channels(O._conv1x1(bf(q["pn.w2"]), hh), q["pn.g2"], q["pn.be2"], 1e-6)).max(-1)[0]
    print(variant, "kernel vs bf16-emulation rel:", float((pooled.cpu() - emu).norm() / emu.norm()), " emulation vs fp32 oracle:", float((emu - ref).norm() / ref.norm()))
    err = (pooled.cpu() - ref)
    neg = q["pn.g2"] < 0
    print(variant, "rel", float(err.norm() / ref.norm()), "| err on g2<0 channels", float(err[:, neg].abs().max()) if neg.any() else 0,
          "| on g2>=0", float(err[:, ~neg].abs().max()), "| ref max", float(ref.max()), "npos", int((~neg).sum()))
    blk = err.abs().reshape(R, 8, 32).mean(dim=(0, 2))
    print("   mean |err| per 32-channel block:", [round(v, 4) for v in blk.tolist()])
    # fp32 kernel path on the same parameters (sanity of the reference)
    pooled32 = torch.empty(R, 256, device="cuda"); am = torch.empty(R, 256, dtype=torch.int32, device="cuda")
    nb = int(L.pointnet_fwd_f32_workspace(R, NP, 128, 128, 256)); ws = torch.empty(nb, dtype=torch.uint8, device="cuda")
    L.pointnet_fwd_f32(xf, R, N, NP, 8, C, d["pn.w0"], d["pn.b0"], d["pn.w1"], d["pn.g1"], d["pn.be1"], d["pn.w2"], d["pn.g2"], d["pn.be2"], 128, 128, 256, 1e-6, pooled32, am, ws, nb, sp())
    print("   fp32 kernel vs oracle rel:", float((pooled32.cpu() - ref).norm() / ref.norm()))
    worst = err.abs().max(0).values.topk(5)
    print("   worst channels", worst.indices.tolist(), [round(v, 4) for v in worst.values.tolist()], "g2 there", q["pn.g2"][worst.indices].tolist())
