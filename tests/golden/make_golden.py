"""Generate golden vectors by running the REFERENCE's own code (read-only /root/reference) on CPU.

Run in the build container only:   python tests/golden/make_golden.py
Writes tests/golden/*.npz.  The reference has no tests/golden vectors of its own,
so these fixtures -- outputs of pyrl's unmodified SAC / DrQ / PointNet classes built from pyrl's own
config files -- are what pins `oracle/pointnet_sac_oracle.py` and, through it, the CUDA path.

The 1024-wide actor/critic MLPs are shrunk to 64 via config overrides so the fixtures stay small; the
PointNet keeps the config's real widths.  No reference code is copied: the agent is built through
Config.fromfile -> replace_placeholder_with_args -> build_agent exactly as run_rl.py does.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import pointnet_sac_oracle as O  # noqa: E402
from oracle.ref_loader import build_reference_agent, load_reference  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
HIDDEN = 64


class FakeMemory:
    """`memory.sample(n)` -> DictArray, the only replay interface update_parameters uses (sac.py:104)."""

    def __init__(self, ns, batch):
        self.ns, self.batch = ns, batch

    def sample(self, n):
        b = {k: (dict(v) if isinstance(v, dict) else v) for k, v in self.batch.items()}
        b["prev_actions"] = np.zeros_like(b["actions"])
        b["episode_dones"] = b["dones"].copy()
        return self.ns.DictArray(b)


def obs_shape_of(obs):
    return {k: (list(v.shape[1:]) if v.ndim > 2 else int(v.shape[1])) for k, v in obs.items()}


def draw_noise(seed, algo, aug, k, B, N, A, with_actor, rng):
    """Re-draw, with the same seed and in the same order, what the reference draws from the global
    CPU RNG inside update_parameters: jitter(obs), jitter(next_obs) (pcd_aug.py:318), the rotation
    angles (pcd_aug.py:186) or the per-cloud translations (pcd_aug.py:193), eps for a' (distributions.py:117), eps for pi."""
    from torch.distributions.utils import _standard_normal

    torch.manual_seed(seed)
    np.random.seed(seed % (2**32))  # RandomDownSample draws n_drop from numpy's global RNG (pcd_aug.py:245)
    noise = {}
    if algo == "drq":
        for which in ("obs", "next"):
            if aug == "jitter":
                noise[f"jitter_{which}"] = torch.FloatTensor(B * k, 3, N).uniform_(*rng)
            elif aug == "rot":
                noise[f"angle_{which}"] = torch.zeros([B * k, 1]).uniform_(*rng)
            elif aug == "downsample":  # pcd_aug.py:244-251 + array_ops.py:659-673; rng = (drop_ratio, fixed_ratio)
                n_drop = int(N * rng[0]) if rng[1] else np.random.randint(int(N * rng[0]))
                noise[f"keep_{which}"] = torch.rand(1, N).argsort(1)[0, : N - n_drop].clone()
            elif aug == "colorjitter":  # torchvision ColorJitter.get_params: randperm(4), then one uniform per op
                order = torch.randperm(4)
                f = [float(torch.empty(1).uniform_(max(0.0, 1 - m), 1 + m)) for m in rng[:3]]
                f.append(float(torch.empty(1).uniform_(-rng[3], rng[3])))
                noise[f"cj_{which}"] = torch.tensor([float(v) for v in order] + f, dtype=torch.float64)
            elif aug == "shift":  # pcd_aug.py:193, translation_range = [hi, hi, hi]
                noise[f"shift_{which}"] = (torch.rand([B * k, 3]) - 0.5) * 2 * torch.tensor([rng[1]] * 3, dtype=torch.float)
    noise["eps_next"] = _standard_normal((B * k, A), dtype=torch.float32, device=torch.device("cpu"))
    if with_actor:
        noise["eps_pi"] = _standard_normal((B, A), dtype=torch.float32, device=torch.device("cpu"))
    return noise


def flatten(prefix, d, out):
    for k, v in d.items():
        if isinstance(v, dict):
            flatten(f"{prefix}{k}/", v, out)
        else:
            out[f"{prefix}{k}"] = v.detach().cpu().numpy() if torch.is_tensor(v) else np.asarray(v)


def gen_update_fixture(ns, name, cfg_path, algo, aug, aug_rng, B, N, A, n_seg, n_pos, S, dup, n_updates=4):
    batch = O.synthetic_batch(seed=0, B=B, N=N, A=A, n_seg=n_seg, n_pos=n_pos, state_dim=S, duplicate_tail=dup)
    obs_shape = obs_shape_of(batch["obs"])
    D_expr = {"sac": "50", "drq": "128"}[algo]
    in_actor = D_expr + (" + agent_shape" if S else "")
    overrides = {
        "batch_size": B,
        "actor_cfg.nn_cfg.mlp_cfg.mlp_spec": [in_actor, HIDDEN, HIDDEN, "action_shape * 2"],
        "critic_cfg.nn_cfg.mlp_cfg.mlp_spec": [in_actor + " + action_shape", HIDDEN, HIDDEN, 1],
    }
    torch.manual_seed(0)
    agent, cfg = build_reference_agent(ns, cfg_path, obs_shape, A, overrides)
    out = {}
    flatten("batch/", batch, out)
    sd0 = {k: v.clone() for k, v in agent.state_dict().items()}
    flatten("init/", O.params_from_reference_state_dict(sd0), out)
    k = int(getattr(agent, "num_aug", 1)) if algo == "drq" else 1
    mem = FakeMemory(ns, batch)
    for u in range(1, n_updates + 1):
        seed = 1000 + u
        with_actor = u % agent.actor_update_interval == 0
        noise = draw_noise(seed, algo, aug, k, B, N, A, with_actor, aug_rng)
        flatten(f"noise{u}/", noise, out)
        torch.manual_seed(seed)
        np.random.seed(seed % (2**32))
        ret = agent.update_parameters(mem, updates=u)
        for key, val in ret.items():
            out[f"ret{u}/{key}"] = np.float64(val)
        flatten(f"after{u}/", O.params_from_reference_state_dict(agent.state_dict()), out)
    meta = dict(
        algo=algo, aug=aug or "", aug_lo=aug_rng[0] if aug_rng else 0.0, aug_hi=aug_rng[1] if aug_rng else 0.0,
        **({f"cj_{n}": float(v) for n, v in zip("bcsh", aug_rng)} if aug == "colorjitter" else {}),
        B=B, N=N, A=A, n_seg=n_seg, n_pos=n_pos, S=S, num_aug=k, n_updates=n_updates,
        gamma=float(agent.gamma), reward_scale=float(agent.reward_scale),
        target_entropy=float(agent.target_entropy), tau=float(agent.update_coeff["default"]),
        actor_update_interval=int(agent.actor_update_interval), target_update_interval=int(agent.target_update_interval),
    )
    for key, val in meta.items():
        out[f"meta/{key}"] = np.asarray(val)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
    print(name, "ok:", {k: round(v, 5) for k, v in ret.items()})


def gen_checkpoint_fixture(ns, name="ref_checkpoint_drq_small"):
    """A checkpoint written by the REFERENCE's own save_checkpoint (checkpoint_utils.py:238-266) after two updates of the
    tiny DrQ jitter configuration, plus -- for the test's cross-check -- every parameter's Adam moments keyed by the
    parameter's NAME (found by object identity between optimizer.param_groups and named_parameters), and the inputs
    and the reference's returned scalars of a third update taken from the saved state."""
    from pyrl.utils.torch.checkpoint_utils import save_checkpoint

    g = dict(np.load(os.path.join(OUT, "drq_jitter_small.npz")))
    meta = {k.split("/")[1]: g[k].item() for k in g if k.startswith("meta/")}
    B, N, A, S = meta["B"], meta["N"], meta["A"], meta["S"]
    batch = O.synthetic_batch(seed=0, B=B, N=N, A=A, n_seg=meta["n_seg"], n_pos=0, state_dim=S, duplicate_tail=True)
    overrides = {
        "batch_size": B,
        "actor_cfg.nn_cfg.mlp_cfg.mlp_spec": ["128 + agent_shape", HIDDEN, HIDDEN, "action_shape * 2"],
        "critic_cfg.nn_cfg.mlp_cfg.mlp_spec": ["128 + agent_shape + action_shape", HIDDEN, HIDDEN, 1],
    }
    torch.manual_seed(0)
    agent, _ = build_reference_agent(ns, "configs/mfrl/drq/maniskill/pn_jitter.py", obs_shape_of(batch["obs"]), A, overrides)
    mem = FakeMemory(ns, batch)
    for u in (1, 2):
        torch.manual_seed(1000 + u)
        agent.update_parameters(mem, updates=u)
    path = os.path.join(OUT, name + ".ckpt")
    save_checkpoint(agent, path, meta={"updates": 2})
    out = {}
    names = {id(p): n for n, p in agent.named_parameters()}
    for opt_name in ("actor_optim", "critic_optim", "alpha_optim"):
        opt = getattr(agent, opt_name)
        for grp in opt.param_groups:
            for prm in grp["params"]:
                st = opt.state[prm]
                out[f"adam/{opt_name}/{names[id(prm)]}/m"] = st["exp_avg"].detach().numpy().copy()
                out[f"adam/{opt_name}/{names[id(prm)]}/v"] = st["exp_avg_sq"].detach().numpy().copy()
                out[f"adam/{opt_name}/{names[id(prm)]}/step"] = np.float64(float(st["step"]))
    flatten("params/", O.params_from_reference_state_dict(agent.state_dict()), out)
    noise = draw_noise(1003, "drq", "jitter", 2, B, N, A, False, (-0.01, 0.01))
    flatten("noise3/", noise, out)
    torch.manual_seed(1003)
    ret = agent.update_parameters(mem, updates=3)
    for key, val in ret.items():
        out[f"ret3/{key}"] = np.float64(val)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
    print(name, "ok:", os.path.getsize(path), "bytes;", {k: round(v, 5) for k, v in ret.items()})


def gen_pointnet_fixture(ns, name, C_extra, B, N, dup, widths=(128, 128, 256), D=128):
    """PointNet forward through the reference's NETWORK registry (pointnet.py), hooks capture the
    ConvMLP output so pooled values + argmax (h.max(-1)) are recorded too."""
    torch.manual_seed(1)
    n_seg, n_pos = C_extra
    rs = np.random.RandomState(7)
    obs = O.synthetic_obs(rs, B, N, n_seg=n_seg, n_pos=n_pos, duplicate_tail=dup)
    C = 6 + n_seg + n_pos
    net = ns.build_all(dict(type="PointNet", feat_dim=C, mlp_spec=list(widths), out_channels=D, feature_transform=[], ignore_first_ln=True))
    # non-trivial affine LN parameters so gamma/beta are exercised
    with torch.no_grad():
        for pname, prm in net.named_parameters():
            if "norm" in pname or "final_mlp.1" in pname:
                prm.add_(0.1 * torch.randn_like(prm))
    captured = {}
    net.conv.register_forward_hook(lambda m, i, o: captured.__setitem__("h", o.detach().clone()))
    with torch.no_grad():
        feat = net({k: torch.from_numpy(v) for k, v in obs.items()})
    pooled, idx = captured["h"].max(-1)
    sd = {"actor.backbone.visual_nn." + k: v for k, v in net.state_dict().items()}
    out = {}
    flatten("obs/", obs, out)
    prm = {}
    for ours, ref in O.reference_key_map().items():
        if ours.startswith("pn."):
            t = sd[ref].detach().clone().float()
            prm[ours] = t.reshape(t.shape[0], t.shape[1]) if t.ndim == 3 else t
    flatten("params/", prm, out)
    out["feat"], out["pooled"], out["idx"] = feat.numpy(), pooled.numpy(), idx.numpy()
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
    print(name, "ok:", feat.shape, "distinct argmax/cloud:", [len(set(r.tolist())) for r in idx][:4])


def main():
    ns = load_reference()
    torch.set_num_threads(8)
    if "--only-downsample" in sys.argv:  # added after the other fixtures were committed; they are not regenerated
        gen_update_fixture(ns, "drq_downsample_small", "configs/mfrl/drq/maniskill/pn_dropout.py", "drq", "downsample",
                           (0.3, 0), B=5, N=80, A=4, n_seg=2, n_pos=0, S=9, dup=False)
        return
    if "--only-checkpoint" in sys.argv:
        gen_checkpoint_fixture(ns)
        return
    if "--only-colorjitter" in sys.argv:  # added in round 2; the other fixtures are not regenerated
        gen_update_fixture(ns, "drq_colorjitter_small", "configs/mfrl/drq/maniskill/pn_colorjitter.py", "drq", "colorjitter",
                           (0.4, 0.4, 0.4, 0.5), B=5, N=88, A=4, n_seg=1, n_pos=0, S=9, dup=False)
        return
    if "--only-shift" in sys.argv:  # added after the other fixtures were committed; they are not regenerated
        gen_update_fixture(ns, "drq_shift_small", "configs/mfrl/drq/maniskill/pn_shift.py", "drq", "shift", (-0.1, 0.1),
                           B=5, N=72, A=4, n_seg=1, n_pos=0, S=9, dup=False)
        return
    gen_pointnet_fixture(ns, "pointnet_fwd_c7", (1, 0), B=3, N=1200, dup=False)
    gen_pointnet_fixture(ns, "pointnet_fwd_c7_dup", (1, 0), B=3, N=1200, dup=True)
    gen_pointnet_fixture(ns, "pointnet_fwd_c9_dmc", (0, 3), B=2, N=1023, dup=False, widths=(64, 128, 256), D=50)
    gen_update_fixture(ns, "sac_dmc_small", "configs/mfrl/sac/dm_control/pn.py", "sac", None, None,
                       B=8, N=128, A=6, n_seg=0, n_pos=0, S=0, dup=False)
    gen_update_fixture(ns, "drq_jitter_small", "configs/mfrl/drq/maniskill/pn_jitter.py", "drq", "jitter", (-0.01, 0.01),
                       B=6, N=96, A=5, n_seg=1, n_pos=0, S=13, dup=True)
    gen_update_fixture(ns, "drq_rot_small", "configs/mfrl/drq/maniskill/pn_rot.py", "drq", "rot", (-0.15, 0.15),
                       B=4, N=80, A=5, n_seg=3, n_pos=0, S=13, dup=False)
    gen_update_fixture(ns, "drq_shift_small", "configs/mfrl/drq/maniskill/pn_shift.py", "drq", "shift", (-0.1, 0.1),
                       B=5, N=72, A=4, n_seg=1, n_pos=0, S=9, dup=False)
    gen_update_fixture(ns, "drq_downsample_small", "configs/mfrl/drq/maniskill/pn_dropout.py", "drq", "downsample",
                       (0.3, 0), B=5, N=80, A=4, n_seg=2, n_pos=0, S=9, dup=False)


if __name__ == "__main__":
    main()
