"""Synthetic replay batches and random-init weights of the named shapes.

Used by bench.py and the examples: there is no network for datasets or checkpoints, so the benchmark
runs on synthetic point clouds with torch-default (kaiming-uniform) initial weights."""
import math

import numpy as np
import torch


def synthetic_obs(rs, B, N, n_seg=0, n_pos=0, state_dim=0):
    obs = {
        "xyz": rs.uniform(-1, 1, size=(B, 3, N)).astype(np.float32),
        "rgb": rs.randint(0, 256, size=(B, 3, N)).astype(np.uint8),
    }
    if n_pos:
        frame = rs.randint(0, n_pos, size=(B, N))
        obs["pos_encoding"] = (np.arange(n_pos)[None, :, None] == frame[:, None, :]).astype(np.uint8)
    if n_seg:
        obs["seg"] = rs.rand(B, n_seg, N) < 0.5
    if state_dim:
        obs["agent"] = rs.randn(B, state_dim).astype(np.float32)
    return obs


def synthetic_batch(seed, B, N, A, n_seg=0, n_pos=0, state_dim=0):
    """Same layout as `ReplayMemory.sample(B)` of the reference (env/replay_buffer.py:297-322)."""
    rs = np.random.RandomState(seed)
    return {
        "obs": synthetic_obs(rs, B, N, n_seg, n_pos, state_dim),
        "next_obs": synthetic_obs(rs, B, N, n_seg, n_pos, state_dim),
        "actions": rs.uniform(-1, 1, size=(B, A)).astype(np.float32),
        "rewards": rs.randn(B, 1).astype(np.float32),
        "dones": np.zeros((B, 1), dtype=bool),
    }


def init_params(seed, spec, zero_out_logstd=False, alpha=0.1):
    """torch's default Conv1d/Linear init; LayerNorm weight 1, bias 0; log_alpha = ln(alpha) (sac.py:83-84)."""
    g = torch.Generator().manual_seed(seed)

    def lin(out_f, in_f):
        bound = 1.0 / math.sqrt(in_f)
        return (torch.rand(out_f, in_f, generator=g) * 2 - 1) * bound, (torch.rand(out_f, generator=g) * 2 - 1) * bound

    c1, c2, c3 = spec.widths
    D, S, A = spec.out_dim, spec.state_dim, spec.action_dim
    h1, h2 = spec.hidden
    p = {}
    p["pn.w0"], p["pn.b0"] = lin(c1, spec.C)
    p["pn.w1"], _ = lin(c2, c1)
    p["pn.g1"], p["pn.be1"] = torch.ones(c2), torch.zeros(c2)
    p["pn.w2"], _ = lin(c3, c2)
    p["pn.g2"], p["pn.be2"] = torch.ones(c3), torch.zeros(c3)
    p["pn.wf"], p["pn.bf"] = lin(D, c3)
    p["pn.gf"], p["pn.bef"] = torch.ones(D), torch.zeros(D)
    for net, (din, dout) in {"actor": (D + S, 2 * A), "q0": (D + S + A, 1), "q1": (D + S + A, 1)}.items():
        p[f"{net}.w0"], p[f"{net}.b0"] = lin(h1, din)
        p[f"{net}.w1"], p[f"{net}.b1"] = lin(h2, h1)
        p[f"{net}.w2"], p[f"{net}.b2"] = lin(dout, h2)
    if zero_out_logstd:  # mlp.py:78-83
        p["actor.w2"][A:] = (torch.rand(A, h2, generator=g) * 2 - 1) * 1e-3
        p["actor.b2"][A:] = (torch.rand(A, generator=g) * 2 - 1) * 1e-3
    p["log_alpha"] = torch.ones(1) * float(np.log(np.float32(alpha)))
    return p


class SimpleBox:
    """gym.spaces.Box stand-in: the actor head only reads low / high / is_bounded() (actor_critic.py:69-71)."""

    def __init__(self, low, high, shape):
        self.low, self.high, self.shape = np.full(shape, low, np.float32), np.full(shape, high, np.float32), tuple(shape)

    def is_bounded(self):
        return True


def obs_shape_of(obs):
    """env_params['obs_shape'] of a batched observation dict (dict_array.py:364-374: lists for >= 2-D leaves, ints for 1-D)."""
    return {k: (list(v.shape[1:]) if np.ndim(v) > 2 else int(np.shape(v)[1])) for k, v in obs.items()}


def make_agent(config_rel, obs_shape, action_dim, overrides=None, **agent_kwargs):
    """Agent from one of the packaged config files, built exactly like run_rl.py builds the reference's
    (Config.fromfile -> env_params -> replace_placeholder_with_args -> build_agent; run_rl.py:505-543).
    `overrides`: dotted keys under agent_cfg (e.g. {"actor_cfg.nn_cfg.mlp_cfg.mlp_spec": [...]}); `agent_kwargs`:
    top-level agent options (batch_size, precision, use_cuda_graph, seed, ...)."""
    from . import Config, config_path, get_kwargs_from_shape, replace_placeholder_with_args
    from .agents import build_agent

    cfg = Config.fromfile(config_path(config_rel))
    merged = {f"agent_cfg.{k}": v for k, v in {**(overrides or {}), **agent_kwargs}.items()}
    if merged:
        cfg.merge_from_dict(merged)
    cfg.agent_cfg["env_params"] = dict(obs_shape=obs_shape, action_shape=action_dim,
                                       action_space=SimpleBox(-1.0, 1.0, (action_dim,)), is_discrete=False)
    cfg = replace_placeholder_with_args(cfg, **get_kwargs_from_shape(obs_shape, action_dim))
    return build_agent(cfg.agent_cfg)
