"""CPU oracle: fp32 restatement of pyrl's PointNet + SAC/DrQ update path.

TEST INFRASTRUCTURE ONLY.  Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s CPU-baseline /
`--impl reference` legs may import this module, and only as the checker (or as the timed CPU
baseline) -- never as part of the product path.  `pointcloud_rl_b200/` must not import it.

Parity pin: the reference ships no tests or golden vectors for this path, so
this oracle is pinned against outputs of the reference ITSELF, executed in the build container by
`tests/golden/make_golden.py` (through `oracle/ref_loader.py`) and committed under `tests/golden/`.
`tests/test_oracle_golden.py` checks every function below against those fixtures.

Everything is plain PyTorch fp32 on the CPU (this is a floating-point path); gradients come from
torch autograd over the restated forward, the optimiser is restated by hand.

Each function cites the reference code it follows (paths relative to /root/reference).
"""
import math

import numpy as np
import torch
import torch.nn.functional as F

# ----------------------------------------------------------------------------------------------
# Parameter naming.  `state["params"]` is a flat dict of fp32 tensors:
#   pn.w0 [c1,C] pn.b0 [c1] | pn.w1 [c2,c1] pn.g1 pn.be1 [c2] | pn.w2 [c3,c2] pn.g2 pn.be2 [c3]
#   pn.wf [D,c3] pn.bf pn.gf pn.bef [D]
#   {actor,q0,q1,tq0,tq1}.{w0,b0,w1,b1,w2,b2}
#   log_alpha [1]
# ----------------------------------------------------------------------------------------------

PN_KEYS = ["pn.w0", "pn.b0", "pn.w1", "pn.g1", "pn.be1", "pn.w2", "pn.g2", "pn.be2", "pn.wf", "pn.bf", "pn.gf", "pn.bef"]
MLP_KEYS = ["w0", "b0", "w1", "b1", "w2", "b2"]

_PN_REF_NAMES = {
    "pn.w0": "conv.mlp.conv0.weight",
    "pn.b0": "conv.mlp.conv0.bias",
    "pn.w1": "conv.mlp.conv1.weight",
    "pn.g1": "conv.mlp.norm1.weight",
    "pn.be1": "conv.mlp.norm1.bias",
    "pn.w2": "conv.mlp.conv2.weight",
    "pn.g2": "conv.mlp.norm2.weight",
    "pn.be2": "conv.mlp.norm2.bias",
    "pn.wf": "final_mlp.0.weight",
    "pn.bf": "final_mlp.0.bias",
    "pn.gf": "final_mlp.1.weight",
    "pn.bef": "final_mlp.1.bias",
}
_MLP_REF_NAMES = {
    "w0": "linear0.weight",
    "b0": "linear0.bias",
    "w1": "linear1.weight",
    "b1": "linear1.bias",
    "w2": "linear2.weight",
    "b2": "linear2.bias",
}
_NET_REF_PREFIX = {
    "actor": "actor.backbone.final_mlp.mlp.",
    "q0": "critic.values.0.backbone.final_mlp.mlp.",
    "q1": "critic.values.1.backbone.final_mlp.mlp.",
    "tq0": "target_critic.values.0.backbone.final_mlp.mlp.",
    "tq1": "target_critic.values.1.backbone.final_mlp.mlp.",
}


def reference_key_map():
    """oracle name -> reference `agent.state_dict()` key."""
    m = {k: "actor.backbone.visual_nn." + v for k, v in _PN_REF_NAMES.items()}
    for net, prefix in _NET_REF_PREFIX.items():
        for k, v in _MLP_REF_NAMES.items():
            m[f"{net}.{k}"] = prefix + v
    m["log_alpha"] = "log_alpha"
    return m


def params_from_reference_state_dict(sd):
    out = {}
    for ours, ref in reference_key_map().items():
        t = torch.as_tensor(np.asarray(sd[ref])).detach().clone().float()
        if ours in ("pn.w0", "pn.w1", "pn.w2"):
            t = t.reshape(t.shape[0], t.shape[1])  # Conv1d k=1 weight [out,in,1]
        out[ours] = t.contiguous()
    return out


# ----------------------------------------------------------------------------------------------
# PointNet  (pyrl/networks/backbones/pointnet.py)
# ----------------------------------------------------------------------------------------------


def preprocess(obs):
    """PointCloudBase.preprocess, pointnet.py:48-63: cat([xyz, rgb/255 if uint8, pos_encoding, seg], dim=-2)."""
    feats = [torch.as_tensor(obs["xyz"]).float()]
    if "rgb" in obs:
        rgb = torch.as_tensor(obs["rgb"])
        feats.append(rgb / 255.0 if rgb.dtype == torch.uint8 else rgb.float())
    for key in ("pos_encoding", "seg"):
        if key in obs:
            feats.append(torch.as_tensor(obs[key]).float())
    return torch.cat(feats, dim=-2)


def _ln_channels(y, gamma, beta, eps):
    """LN1d (nn_layer.py:209-219): permute to channels-last, nn.LayerNorm over C (biased variance),
    permute back.  Same torch primitive as the reference so duplicated points stay bit-identical."""
    z = F.layer_norm(y.permute(0, 2, 1).contiguous(), (y.shape[1],), gamma, beta, eps)
    return z.permute(0, 2, 1).contiguous()


def _conv1x1(w, x, b=None):
    """nn.Conv1d(kernel_size=1) -- the per-point shared linear map (block_utils.py:87)."""
    return F.conv1d(x, w[:, :, None], b)


def pointnet_point_features(p, x, ln_eps=1e-6):
    """ConvMLP with ignore_first_ln (mlp.py:43-56, block_utils.py:60-100): [B,C,N] -> [B,c3,N]."""
    h = torch.relu(_conv1x1(p["pn.w0"], x, p["pn.b0"]))
    h = torch.relu(_ln_channels(_conv1x1(p["pn.w1"], h), p["pn.g1"], p["pn.be1"], ln_eps))
    h = torch.relu(_ln_channels(_conv1x1(p["pn.w2"], h), p["pn.g2"], p["pn.be2"], ln_eps))
    return h


def pointnet_forward(p, x, ln_eps=1e-6, return_pool=False, idx_override=None):
    """PointNet.forward, pointnet.py:112-157 with feature_transform=[]: per-point MLP, max over
    points (ties -> smallest index, torch semantics), Linear + LayerNorm(eps=1e-5) (pointnet.py:110).

    idx_override [B,c3] (tests of the reduced-precision tier only): pool the points a kernel selected instead of
    the fp32 argmax.  max-pool routes each channel's gradient to ONE point, so a near-tie resolved differently under
    bf16 rounding moves that channel's whole gradient; with the selection pinned, the remaining difference is
    arithmetic error, which is what the tolerance is about (the selection itself is checked by the near-tie rule)."""
    h = pointnet_point_features(p, x, ln_eps)
    if idx_override is None:
        pooled, idx = h.max(dim=-1)
    else:
        idx = idx_override.long()
        pooled = torch.gather(h, 2, idx[..., None])[..., 0]
    z = pooled @ p["pn.wf"].t() + p["pn.bf"]
    out = F.layer_norm(z, (z.shape[-1],), p["pn.gf"], p["pn.bef"], 1e-5)
    if return_pool:
        return out, pooled, idx
    return out


# ----------------------------------------------------------------------------------------------
# MLP heads, policy head
# ----------------------------------------------------------------------------------------------


def mlp3(p, net, x):
    """LinearMLP with norm_cfg=None, inactivated_output=True (mlp.py:98-100): Linear-ReLU-Linear-ReLU-Linear."""
    h = torch.relu(x @ p[f"{net}.w0"].t() + p[f"{net}.b0"])
    h = torch.relu(h @ p[f"{net}.w1"].t() + p[f"{net}.b1"])
    return h @ p[f"{net}.w2"].t() + p[f"{net}.b2"]


def tanh_gaussian(out, eps, log_std_bound=(-10.0, 2.0), scale=1.0, bias=0.0, epsilon=1e-6):
    """TanhGaussianHead 'max-entropy' mode.  gaussian.py:36,49 (chunk, clamp, exp);
    distributions.py:89,116-119 (u = mu + sigma*eps; logp -= log(scale*(1-tanh(u)^2)+1e-6));
    regression_base.py:70-72 returns (sample, -logp[..., None])."""
    mean, log_std = out.chunk(2, dim=-1)
    std = torch.clamp(log_std, min=log_std_bound[0], max=log_std_bound[1]).exp()
    u = mean + std * eps
    # torch.distributions.Normal.log_prob
    logp = -((u - mean) ** 2) / (2 * std**2) - std.log() - math.log(math.sqrt(2 * math.pi))
    t = torch.tanh(u)
    logp = logp - torch.log(scale * (1 - t.pow(2)) + epsilon)
    return t * scale + bias, -logp.sum(-1, keepdim=True)


# ----------------------------------------------------------------------------------------------
# Augmentations (pyrl/utils/augmentations/pcd_aug.py)
# ----------------------------------------------------------------------------------------------


def aug_jitter(xyz, noise):
    """RandomJitterPoints.process_single, pcd_aug.py:316-322: xyz + U(lo,hi) noise (noise injected)."""
    return xyz + noise


def aug_rot_z(xyz, angle):
    """GlobalRotScaleTrans (rot only), pcd_aug.py:186-187 + ops.py:171-183 + apply_rot_trans einsum
    'bin,bji->bjn': x' = R x, R = [[c,-s,0],[s,c,0],[0,0,1]]; angle [B,1]."""
    c, s = torch.cos(angle)[:, 0], torch.sin(angle)[:, 0]
    rot = torch.zeros(xyz.shape[0], 3, 3)
    rot[:, 2, 2] = 1
    rot[:, 0, 0] = c
    rot[:, 1, 1] = c
    rot[:, 0, 1] = -s
    rot[:, 1, 0] = s
    return torch.einsum("bin,bji->bjn", xyz, rot)


def aug_shift(xyz, shift):
    """GlobalRotScaleTrans (translation only, shift_height=True), pcd_aug.py:192-197 + apply_rot_trans :84-123:
    xyz [B,3,N] + shift [B,3][..., None]."""
    return xyz + shift[:, :, None]


def aug_downsample(obs, index):
    """RandomDownSample.process_single, pcd_aug.py:240-257: every point-cloud key keeps the same `index` subset of its
    points (`DictArray(data).slice(index, -1)`); index [N'] int64 (the argsort-of-rand draw is injected)."""
    return {k: (v[..., index] if k in ("xyz", "rgb", "seg", "pos_encoding") else v) for k, v in obs.items()}


def aug_colorjitter(rgb, params):
    """ColorJitterPoints.process_single, pcd_aug.py:291-295: torchvision ColorJitter.forward on rgb[:, :, None, :]
    (uint8 [B',3,1,N]).  torchvision is a third-party dependency of the reference (environment.yml); its forward is
    restated here with the draws injected: params = [order0..3, brightness, contrast, saturation, hue] -- the op order
    (a permutation of 0 brightness, 1 contrast, 2 saturation, 3 hue) and one factor per op, shared by the whole batch --
    calling torchvision's own adjust_* functions."""
    import torchvision.transforms.functional as TF

    img = torch.as_tensor(rgb)[:, :, None, :]
    order = [int(v) for v in params[:4]]
    b, c, s, h = (float(v) for v in params[4:8])
    for fn_id in order:
        if fn_id == 0:
            img = TF.adjust_brightness(img, b)
        elif fn_id == 1:
            img = TF.adjust_contrast(img, c)
        elif fn_id == 2:
            img = TF.adjust_saturation(img, s)
        elif fn_id == 3:
            img = TF.adjust_hue(img, h)
    return img.squeeze(-2)


def _apply_aug(obs, hp, noise, which):
    obs = dict(obs)
    kind = hp.get("aug", None)
    if kind == "jitter":
        obs["xyz"] = aug_jitter(obs["xyz"], noise[f"jitter_{which}"])
    elif kind == "rot":
        obs["xyz"] = aug_rot_z(obs["xyz"], noise[f"angle_{which}"])
    elif kind == "shift":
        obs["xyz"] = aug_shift(obs["xyz"], noise[f"shift_{which}"])
    elif kind == "downsample":
        obs = aug_downsample(obs, noise[f"keep_{which}"].long())
    elif kind == "colorjitter":
        obs["rgb"] = aug_colorjitter(obs["rgb"], noise[f"cj_{which}"])
    elif kind is not None:
        raise ValueError(kind)
    return obs


def _repeat_obs(obs, k):
    """GDict.repeat(k, axis=0) on tensors == repeat_interleave (array_ops.py:106-121)."""
    return {key: torch.repeat_interleave(torch.as_tensor(v), k, dim=0) for key, v in obs.items()}


# ----------------------------------------------------------------------------------------------
# Optimiser / target update
# ----------------------------------------------------------------------------------------------


def adam_step(p, g, st, lr=1e-3, betas=(0.9, 0.999), eps=1e-8):
    """torch.optim.Adam single-tensor semantics (no weight decay, no amsgrad): eps added after the
    bias-corrected sqrt(v).  Reference builds it via optimizer_utils.py:31-64."""
    st["step"] += 1
    t = st["step"]
    st["m"].mul_(betas[0]).add_(g, alpha=1 - betas[0])
    st["v"].mul_(betas[1]).addcmul_(g, g, value=1 - betas[1])
    bc1 = 1 - betas[0] ** t
    bc2 = 1 - betas[1] ** t
    denom = (st["v"].sqrt() / math.sqrt(bc2)).add_(eps)
    p.addcdiv_(st["m"], denom, value=-(lr / bc1))


def new_adam_state(params, keys):
    return {k: {"step": 0, "m": torch.zeros_like(params[k]), "v": torch.zeros_like(params[k])} for k in keys}


def critic_keys():
    return PN_KEYS + [f"q0.{k}" for k in MLP_KEYS] + [f"q1.{k}" for k in MLP_KEYS]


def actor_keys():
    return [f"actor.{k}" for k in MLP_KEYS]


def new_state(params):
    params = {k: v.detach().clone().float() for k, v in params.items()}
    return {
        "params": params,
        "critic_optim": new_adam_state(params, critic_keys()),
        "actor_optim": new_adam_state(params, actor_keys()),
        "alpha_optim": new_adam_state(params, ["log_alpha"]),
        "alpha": float(params["log_alpha"].exp().item()),  # sac.py:99-100
    }


def grad_norm(grads):
    """ExtendedModuleBase.grad_norm, module_utils.py:40-45: L2 norm of per-tensor L2 norms."""
    return float(torch.norm(torch.stack([torch.norm(g, 2) for g in grads]), 2).item())


# ----------------------------------------------------------------------------------------------
# The update step  (sac.py:103-214, drq.py:46-165; ordered spec in SURVEY.md Appendix B)
# ----------------------------------------------------------------------------------------------

DEFAULT_HP = dict(
    algo="sac",
    gamma=0.99,
    reward_scale=1.0,
    num_aug=1,
    aug=None,
    tau=0.01,
    actor_update_interval=2,
    target_update_interval=2,
    lr=1e-3,
    alpha_lr=1e-3,
    alpha_betas=(0.5, 0.999),
    log_std_bound=(-10.0, 2.0),
    head_scale=1.0,
    head_bias=0.0,
    target_entropy=None,
    ignore_dones=False,
    automatic_alpha_tuning=True,
)


def _obs_split(obs):
    obs = {k: torch.as_tensor(v) for k, v in obs.items()}
    robot = None
    for key in ("state", "agent"):  # visuomotor.py:87-91
        if key in obs:
            robot = obs.pop(key).float()
    return obs, robot


def _cat(*xs):
    return torch.cat([x for x in xs if x is not None], dim=-1)


def update(state, batch, updates, hp, noise, capture=None, idx_override=None):
    """One SAC / DrQ `update_parameters(memory, updates)` on an already-sampled batch.

    batch: dict(obs=dict, next_obs=dict, actions [B,A], rewards [B,1], dones [B,1]) of numpy / torch.
    noise: injected randomness -- jitter_obs / jitter_next [B*num_aug,3,N] (or angle_obs / angle_next
           [B*num_aug,1], or shift_obs / shift_next [B*num_aug,3], or keep_obs / keep_next [N'] kept point indices, or
           cj_obs / cj_next [8] colour-jitter order + factors), eps_next [B*num_aug,A], eps_pi [B,A].
    Mutates `state` in place, returns the reference's scalar dict.  `capture` (a dict) receives
    intermediates for the parity tests.
    """
    hp = dict(DEFAULT_HP, **hp)
    p = state["params"]
    pre = hp["algo"]
    k = hp["num_aug"] if pre == "drq" else 1
    cap = capture if capture is not None else {}

    actions = torch.as_tensor(batch["actions"]).float()
    rewards = torch.as_tensor(batch["rewards"]).float()
    dones = torch.as_tensor(batch["dones"]).float()
    B = actions.shape[0]
    A = actions.shape[1]
    target_entropy = hp["target_entropy"] if hp["target_entropy"] is not None else -float(A)  # sac.py:96

    obs, robot = _obs_split(batch["obs"])
    nobs, nrobot = _obs_split(batch["next_obs"])
    if pre == "drq":  # drq.py:58-63
        obs = _apply_aug(_repeat_obs(obs, k), hp, noise, "obs")
        nobs = _apply_aug(_repeat_obs(nobs, k), hp, noise, "next")
        robot = None if robot is None else torch.repeat_interleave(robot, k, 0)
        nrobot = None if nrobot is None else torch.repeat_interleave(nrobot, k, 0)
        actions_k = torch.repeat_interleave(actions, k, 0)
        rewards_k = torch.repeat_interleave(rewards, k, 0)
        dones_k = torch.repeat_interleave(dones, k, 0)
    else:
        actions_k, rewards_k, dones_k = actions, rewards, dones
    x_obs = preprocess(obs)
    x_next = preprocess(nobs)
    cap["x_obs"], cap["x_next"] = x_obs, x_next
    alpha = state["alpha"]
    head = dict(log_std_bound=hp["log_std_bound"], scale=hp["head_scale"], bias=hp["head_bias"])

    # ---- target (no grad): sac.py:108-134 / drq.py:71-87
    with torch.no_grad():
        f_next = pointnet_forward(p, x_next)
        a_next, nlogp_next = tanh_gaussian(mlp3(p, "actor", _cat(f_next, nrobot)), noise["eps_next"], **head)
        q_in = _cat(f_next, nrobot, a_next)
        q_next = torch.cat([mlp3(p, "tq0", q_in), mlp3(p, "tq1", q_in)], dim=-1)
        v = q_next.min(dim=-1, keepdim=True).values + alpha * nlogp_next
        not_done = 1.0 if hp["ignore_dones"] else (1 - dones_k)
        if pre == "drq":
            y = rewards_k + not_done * hp["gamma"] * v  # drq.py:81 (no reward_scale)
            y = y.reshape(B, k).mean(1, keepdim=True)  # drq.py:84
            y = torch.repeat_interleave(y, k, dim=0)
        else:
            y = rewards_k * hp["reward_scale"] + not_done * hp["gamma"] * v  # sac.py:133
        q_target = y.repeat(1, 2)
        cap.update(f_next=f_next, a_next=a_next, nlogp_next=nlogp_next, q_next=q_next, q_target=q_target)

    # ---- critic step: sac.py:136-148 / drq.py:89-101
    ck = critic_keys()
    leaves = {name: p[name].detach().clone().requires_grad_(True) for name in ck}
    f, pooled, idx = pointnet_forward(leaves, x_obs, return_pool=True, idx_override=idx_override)
    q_in = _cat(f, robot, actions_k)
    q = torch.cat([mlp3(leaves, "q0", q_in), mlp3(leaves, "q1", q_in)], dim=-1)
    critic_loss = F.mse_loss(q, q_target) * 2
    grads = torch.autograd.grad(critic_loss, [leaves[name] for name in ck])
    cap.update(f_obs=f.detach(), pooled_obs=pooled.detach(), idx_obs=idx, q=q.detach(), critic_grads=dict(zip(ck, grads)))
    if idx_override is not None:  # every candidate's true post-ReLU feature, for the near-tie check of the selection
        with torch.no_grad():
            cap["h_obs_max"] = pointnet_point_features(p, x_obs).max(dim=-1)[0]
    ret = {
        f"{pre}/critic_loss": float(critic_loss.item()),
        f"{pre}/max_critic_abs_err": float((q - q_target).abs().max().item()),
        f"{pre}/alpha": alpha,
        f"{pre}/q": float(q.min(dim=-1).values.mean().item()),
        f"{pre}/q_target": float(q_target.mean().item()),
        f"{pre}/target_entropy": target_entropy,
        f"{pre}/critic_grad": grad_norm(grads),
        f"{pre}/grad_steps": 1,
    }
    for name, g in zip(ck, grads):
        adam_step(p[name], g, state["critic_optim"][name], lr=hp["lr"])

    # ---- actor + alpha step: sac.py:161-205 / drq.py:114-155
    if updates % hp["actor_update_interval"] == 0:
        if pre == "drq":
            x_pi = x_obs.reshape(B, k, *x_obs.shape[1:])[:, 0]  # first augmentation, drq.py:115
            robot_pi = None if robot is None else robot.reshape(B, k, -1)[:, 0]
        else:
            x_pi, robot_pi = x_obs, robot
        with torch.no_grad():  # detach_actor_feature=True -> visuomotor.py:116-117
            f_pi = pointnet_forward(p, x_pi)  # post-critic-step PointNet weights
        ak = actor_keys()
        aleaves = {name: p[name].detach().clone().requires_grad_(True) for name in ak}
        pi, nlogp = tanh_gaussian(mlp3(aleaves, "actor", _cat(f_pi, robot_pi)), noise["eps_pi"], **head)
        entropy = nlogp.mean()
        q_in = _cat(f_pi, robot_pi, pi)
        q_pi = torch.cat([mlp3(p, "q0", q_in), mlp3(p, "q1", q_in)], dim=-1).min(dim=-1, keepdim=True).values
        actor_loss = -(q_pi.mean() + alpha * entropy)
        agrads = torch.autograd.grad(actor_loss, [aleaves[name] for name in ak])
        cap.update(f_pi=f_pi, pi=pi.detach(), nlogp_pi=nlogp.detach(), actor_grads=dict(zip(ak, agrads)))
        for name, g in zip(ak, agrads):
            adam_step(p[name], g, state["actor_optim"][name], lr=hp["lr"])
        ret[f"{pre}/actor_loss"] = float(actor_loss.item())
        ret[f"{pre}/entropy"] = float(entropy.item())
        ret[f"{pre}/actor_grad"] = grad_norm(agrads)
        if hp["automatic_alpha_tuning"]:
            # alpha_loss = exp(log_alpha) * (entropy - target_entropy).detach()   sac.py:190-195
            la = p["log_alpha"]
            coef = float(entropy.item()) - target_entropy
            alpha_loss = float((la.exp() * coef).item())
            g = la.exp() * coef
            adam_step(la, g, state["alpha_optim"]["log_alpha"], lr=hp["alpha_lr"], betas=hp["alpha_betas"])
            state["alpha"] = float(la.exp().item())
        else:
            alpha_loss = 0.0
        ret[f"{pre}/alpha_loss"] = alpha_loss

    # ---- Polyak: sac.py:207-208, ops.py:60-90 (shared PointNet skipped by the id() guard)
    if updates % hp["target_update_interval"] == 0:
        tau = hp["tau"]
        for h in (0, 1):
            for key in MLP_KEYS:
                t, s = p[f"tq{h}.{key}"], p[f"q{h}.{key}"]
                t.copy_(t * (1.0 - tau) + s * tau)
    return ret


# ----------------------------------------------------------------------------------------------
# Synthetic replay batch -- shared by tests and bench so every arm sees the same bytes
# ----------------------------------------------------------------------------------------------


def synthetic_obs(rs, B, N, n_seg=0, n_pos=0, state_dim=0, duplicate_tail=False):
    obs = {
        "xyz": rs.uniform(-1, 1, size=(B, 3, N)).astype(np.float32),
        "rgb": rs.randint(0, 256, size=(B, 3, N)).astype(np.uint8),
    }
    if n_pos:
        frame = rs.randint(0, n_pos, size=(B, N))
        obs["pos_encoding"] = (np.arange(n_pos)[None, :, None] == frame[:, None, :]).astype(np.uint8)
    if n_seg:
        obs["seg"] = rs.rand(B, n_seg, N) < 0.5
    if duplicate_tail:  # ManiSkill pads clouds by repeating points (observation_process.py:63-65)
        q = N - N // 4
        for key in ("xyz", "rgb", "pos_encoding", "seg"):
            if key in obs:
                obs[key][..., q:] = obs[key][..., : N - q]
    if state_dim:
        obs["agent"] = rs.randn(B, state_dim).astype(np.float32)
    return obs


def synthetic_batch(seed, B, N, A, n_seg=0, n_pos=0, state_dim=0, duplicate_tail=False):
    rs = np.random.RandomState(seed)
    return {
        "obs": synthetic_obs(rs, B, N, n_seg, n_pos, state_dim, duplicate_tail),
        "next_obs": synthetic_obs(rs, B, N, n_seg, n_pos, state_dim, duplicate_tail),
        "actions": rs.uniform(-1, 1, size=(B, A)).astype(np.float32),
        "rewards": rs.randn(B, 1).astype(np.float32),
        "dones": np.zeros((B, 1), dtype=bool),
    }


def init_params(seed, C, widths, D, state_dim, A, hidden=1024, zero_out_logstd=False):
    """Random-init weights with torch's default Conv1d/Linear init (kaiming_uniform(a=sqrt(5))), used
    when no reference fixture is available (full-size parity on the GPU box, bench)."""
    g = torch.Generator().manual_seed(seed)

    def lin(out_f, in_f):
        bound = 1.0 / math.sqrt(in_f)
        w = (torch.rand(out_f, in_f, generator=g) * 2 - 1) * bound
        b = (torch.rand(out_f, generator=g) * 2 - 1) * bound
        return w, b

    c1, c2, c3 = widths
    p = {}
    p["pn.w0"], p["pn.b0"] = lin(c1, C)
    p["pn.w1"], _ = lin(c2, c1)
    p["pn.g1"], p["pn.be1"] = torch.ones(c2), torch.zeros(c2)
    p["pn.w2"], _ = lin(c3, c2)
    p["pn.g2"], p["pn.be2"] = torch.ones(c3), torch.zeros(c3)
    p["pn.wf"], p["pn.bf"] = lin(D, c3)
    p["pn.gf"], p["pn.bef"] = torch.ones(D), torch.zeros(D)
    dims = {"actor": (D + state_dim, 2 * A), "q0": (D + state_dim + A, 1), "q1": (D + state_dim + A, 1)}
    for net, (din, dout) in dims.items():
        p[f"{net}.w0"], p[f"{net}.b0"] = lin(hidden, din)
        p[f"{net}.w1"], p[f"{net}.b1"] = lin(hidden, hidden)
        p[f"{net}.w2"], p[f"{net}.b2"] = lin(dout, hidden)
    if zero_out_logstd:  # mlp.py:78-83
        p["actor.w2"][A:] = (torch.rand(A, hidden, generator=g) * 2 - 1) * 1e-3
        p["actor.b2"][A:] = (torch.rand(A, generator=g) * 2 - 1) * 1e-3
    for h in (0, 1):
        for key in MLP_KEYS:
            p[f"tq{h}.{key}"] = p[f"q{h}.{key}"].clone()  # hard_update, builder.py:43
    p["log_alpha"] = torch.ones(1) * float(np.log(np.float32(0.1)))
    return p
