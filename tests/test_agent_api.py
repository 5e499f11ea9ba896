"""The drop-in boundary: agents built through the registry from the package's config files expose the
reference's module tree / state_dict keys (CPU), and `update_parameters(memory, updates)` reproduces the
reference's returned dict and weights (GPU, golden vectors)."""
import json
import os

import numpy as np
import pytest
import torch

from pointcloud_rl_b200 import Config, config_path, get_kwargs_from_shape, replace_placeholder_with_args
from tests.conftest import GOLDEN, load_golden


class Box:  # gym.spaces.Box stand-in: the actor only reads low/high/is_bounded (actor_critic.py:69-71)
    def __init__(self, low, high, shape):
        self.low, self.high, self.shape = np.full(shape, low, np.float32), np.full(shape, high, np.float32), shape

    def is_bounded(self):
        return True


def make_agent(rel, obs_shape, A, hidden=None, **overrides):
    from pointcloud_rl_b200.agents import build_agent

    cfg = Config.fromfile(config_path(rel))
    if hidden:
        D = cfg.agent_cfg.actor_cfg.nn_cfg.visual_nn_cfg.out_channels
        state = " + agent_shape" if "agent" in obs_shape else ""
        cfg.agent_cfg.actor_cfg.nn_cfg.mlp_cfg.mlp_spec = [f"{D}{state}", hidden, hidden, "action_shape * 2"]
        cfg.agent_cfg.critic_cfg.nn_cfg.mlp_cfg.mlp_spec = [f"{D}{state} + action_shape", hidden, hidden, 1]
    cfg.merge_from_dict({f"agent_cfg.{k}": v for k, v in overrides.items()})
    cfg.agent_cfg["env_params"] = dict(obs_shape=obs_shape, action_shape=A, action_space=Box(-1.0, 1.0, (A,)), is_discrete=False)
    cfg = replace_placeholder_with_args(cfg, **get_kwargs_from_shape(obs_shape, A))
    return build_agent(cfg.agent_cfg)


def test_state_dict_keys_match_reference():
    obs_shape = {"xyz": [3, 64], "rgb": [3, 64], "seg": [1, 64], "agent": 13}
    agent = make_agent("mfrl/drq/maniskill/pn_jitter.py", obs_shape, 5, batch_size=4)
    ours = {k: tuple(v.shape) for k, v in agent.state_dict().items()}
    ref = {k: tuple(v) for k, v in json.load(open(os.path.join(GOLDEN, "state_dict_keys_drq_maniskill.json"))).items()}
    assert ours == ref
    # object identity of the shared backbone (builder.py:28-45,60-66)
    pn = agent.actor.backbone.visual_nn
    assert all(v.backbone.visual_nn is pn for v in agent.critic.values)
    assert all(v.backbone.visual_nn is pn for v in agent.target_critic.values)
    assert agent.target_entropy == -5 and abs(agent.alpha - 0.1) < 1e-6
    assert agent.num_aug == 2 and agent._aug == ("jitter", -0.01, 0.01)
    assert agent.critic.num_trainable_parameters == sum(p.numel() for p in set(agent.critic.parameters()))


def test_unsupported_options_raise():
    obs_shape = {"xyz": [3, 64], "rgb": [3, 64]}
    with pytest.raises(NotImplementedError):
        make_agent("mfrl/sac/dm_control/pn.py", obs_shape, 6, shared_backbone=False)
    from pointcloud_rl_b200.networks import build_all

    with pytest.raises(NotImplementedError):
        build_all(dict(type="PointNet", feat_dim=6, mlp_spec=[64, 128, 256], out_channels=50, feature_transform=[1]))
    from pointcloud_rl_b200.augmentations import build_data_augmentations

    shift = dict(type="GlobalRotScaleTrans", main_key="xyz", req_keys=["xyz"], rot_range=None, scale_ratio_range=None,
                 translation_range=[0.04, 0, 0.04], shift_height=True)
    assert build_data_augmentations(shift)[0].params() == ("shift", -0.04, 0.04, 0b101)  # pn_shift.py is supported
    for bad in (dict(shift, shift_height=False),                 # the reference zeroes the last cloud's shift there
                dict(shift, translation_range=[0.1, 0.2, 0.1]),  # unequal per-axis ranges
                dict(shift, scale_ratio_range=[0.95, 1.05]),
                dict(shift, rot_range=[-0.1, 0.1])):
        with pytest.raises(NotImplementedError):
            build_data_augmentations(bad)


def test_update_needs_cuda_no_cpu_fallback():
    if torch.cuda.is_available():
        pytest.skip("CPU-only check")
    from pointcloud_rl_b200.data import FixedBatchMemory
    from pointcloud_rl_b200.synthetic import synthetic_batch

    agent = make_agent("mfrl/sac/dm_control/pn.py", {"xyz": [3, 32], "rgb": [3, 32]}, 6, hidden=32, batch_size=2)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        agent.update_parameters(FixedBatchMemory(synthetic_batch(0, 2, 32, 6)), 1)


# ------------------------------------------------------------------------------------------ GPU
def _load_golden_into(agent, init):
    from oracle.pointnet_sac_oracle import reference_key_map

    sd = agent.state_dict()
    for ours, ref in reference_key_map().items():
        t = torch.from_numpy(np.asarray(init[ours]))
        sd[ref] = t.reshape(sd[ref].shape)
    for h in (0, 1):  # the critic heads' backbones alias the actor's: fill the duplicate keys too
        for k in list(sd):
            if k.startswith("actor.backbone.visual_nn."):
                for pre in (f"critic.values.{h}", f"target_critic.values.{h}"):
                    sd[k.replace("actor", pre, 1)] = sd[k]
    agent.load_state_dict(sd)


@pytest.mark.gpu
@pytest.mark.parametrize("name,rel", [("sac_dmc_small", "mfrl/sac/dm_control/pn.py"),
                                      ("drq_jitter_small", "mfrl/drq/maniskill/pn_jitter.py")])
def test_update_parameters_through_registry(name, rel):
    from pointcloud_rl_b200.data import FixedBatchMemory

    g = load_golden(name)
    m = {k: v.item() for k, v in g["meta"].items()}
    batch = {k: (dict(v) if isinstance(v, dict) else v) for k, v in g["batch"].items()}
    obs_shape = {k: (list(v.shape[1:]) if v.ndim > 2 else int(v.shape[1])) for k, v in batch["obs"].items()}
    agent = make_agent(rel, obs_shape, m["A"], hidden=64, batch_size=m["B"], precision="fp32", use_cuda_graph=False).to("cuda")
    _load_golden_into(agent, g["init"])
    eng = agent._ensure_engine(batch)
    mem = FixedBatchMemory(batch)
    for u in (1, 2):
        noise = {k: torch.from_numpy(v).cuda() for k, v in g[f"noise{u}"].items()}
        # the public call draws its own (Philox) randomness; parity needs the reference's draws injected
        eng.upload_batch(batch)
        eng.update(u, noise)
        ret = eng.read_scalars(u)
        ref = {f"{a}/{b}": float(v) for a, sub in g[f"ret{u}"].items() for b, v in sub.items()}
        assert set(ret) == set(ref)
        for key, val in ref.items():
            assert ret[key] == pytest.approx(val, rel=1e-3, abs=1e-4), (u, key)
    # module parameters alias the engine's flat buffer: the nn.Module view moved with the fused Adam
    w_mod = agent.actor.backbone.visual_nn.conv.mlp.conv1.weight
    assert w_mod.data_ptr() == eng.p["pn.w1"].data_ptr()
    assert torch.allclose(w_mod[..., 0].cpu(), torch.from_numpy(g["after2"]["pn.w1"]), atol=1e-4)
    # and the public call itself (own randomness) returns the reference's keys with finite values
    out = agent.update_parameters(mem, 4)
    assert set(out) == set(ref) and all(np.isfinite(v) for v in out.values())


@pytest.mark.gpu
def test_rollout_forward_and_modes():
    obs_shape = {"xyz": [3, 200], "rgb": [3, 200], "seg": [1, 200], "agent": 13}
    agent = make_agent("mfrl/drq/maniskill/pn_jitter.py", obs_shape, 5, hidden=64, batch_size=4, precision="fp32").to("cuda")
    from oracle import pointnet_sac_oracle as O

    rs = np.random.RandomState(0)
    obs = O.synthetic_obs(rs, 3, 200, n_seg=1, state_dim=13)
    mean = agent(obs, mode="eval")
    assert mean.shape == (3, 5) and float(mean.abs().max()) <= 1.0
    a, nlp = agent(obs, mode="max-entropy")
    assert a.shape == (3, 5) and nlp.shape == (3, 1)
    # against the oracle with the agent's own weights
    sd = {k: v.detach().cpu() for k, v in agent.state_dict().items()}
    p = O.params_from_reference_state_dict(sd)
    x = O.preprocess({k: torch.from_numpy(v) for k, v in obs.items() if k != "agent"})
    f = O.pointnet_forward(p, x)
    out = O.mlp3(p, "actor", torch.cat([f, torch.from_numpy(obs["agent"])], -1))
    ref_mean = torch.tanh(out[:, :5])
    assert torch.allclose(mean.cpu(), ref_mean, atol=1e-4)
    feat = agent.actor.backbone.visual_nn({k: v for k, v in obs.items() if k != "agent"})
    assert torch.allclose(feat.cpu(), f, atol=1e-4)


@pytest.mark.gpu
def test_device_replay_ring_matches_a_host_ring():
    """transitions live in HBM, `sample` draws the reference sampler's indices on the host and
    gathers on the device straight into the engine's batch buffer (bit-exact data movement)."""
    from oracle import pointnet_sac_oracle as O
    from pointcloud_rl_b200.replay import DeviceReplayMemory

    B, N, A, S, cap = 6, 96, 5, 13, 37
    obs_shape = {"xyz": [3, N], "rgb": [3, N], "seg": [1, N], "agent": S}
    agent = make_agent("mfrl/drq/maniskill/pn_jitter.py", obs_shape, A, hidden=64, batch_size=B, precision="fp32").to("cuda")
    ring = DeviceReplayMemory(cap, device="cuda", seed=5)
    host = None  # numpy mirror with the reference ring's overwrite order (replay_buffer.py:206-231)
    pos = count = 0
    for chunk, n in enumerate((20, 1, 25, 9)):  # 55 transitions through a 37-slot ring: wraps once
        batch = O.synthetic_batch(seed=10 + chunk, B=n, N=N, A=A, n_seg=1, n_pos=0, state_dim=S)
        flat = {f"{w}/{k}": np.asarray(v) for w in ("obs", "next_obs") for k, v in batch[w].items()}
        flat.update({k: np.asarray(batch[k]) for k in ("actions", "rewards", "dones")})
        if host is None:
            host = {k: np.zeros((cap,) + v.shape[1:], v.dtype) for k, v in flat.items()}
        if n == 1:
            ring.push({k: ({kk: vv[0] for kk, vv in v.items()} if isinstance(v, dict) else np.asarray(v)[0]) for k, v in batch.items()})
        else:
            ring.push_batch(batch)
        for i in range(n):
            for k, v in flat.items():
                host[k][pos] = v[i]
            pos, count = (pos + 1) % cap, count + 1
    assert len(ring) == cap and ring.position == pos
    twin = np.random.RandomState(5)
    got = ring.sample(B)
    idx = twin.randint(0, cap, size=B)
    assert np.array_equal(got.index, idx)
    back = got.to_host()
    for k, v in host.items():
        node = back
        for part in k.split("/"):
            node = node[part]
        assert np.array_equal(node.astype(v.dtype).reshape(v[idx].shape), v[idx]), k
    # through the public call: the engine's batch buffer holds exactly the sampled transitions
    out = agent.update_parameters(ring, 2)
    assert all(np.isfinite(v) for v in out.values())
    idx = twin.randint(0, cap, size=B)
    eng = agent.engine
    assert np.array_equal(eng.raw["obs"]["xyz"].cpu().numpy(), host["obs/xyz"][idx])
    assert np.array_equal(eng.raw["next_obs"]["rgb"].cpu().numpy(), host["next_obs/rgb"][idx])
    assert np.array_equal(eng.raw["obs"]["seg"].cpu().numpy().astype(bool), host["obs/seg"][idx].astype(bool))
    assert np.array_equal(eng.raw["obs"]["state"].cpu().numpy(), host["obs/agent"][idx])
    assert np.array_equal(eng.raw["actions"].cpu().numpy(), host["actions"][idx])
    assert np.array_equal(eng.raw["rewards"].cpu().numpy(), host["rewards"][idx].reshape(-1))


@pytest.mark.gpu
@pytest.mark.parametrize("rel", ["mfrl/drq/maniskill/pn_shift.py", "mfrl/drq/maniskill/pn_dropout.py",
                                 "mfrl/drq/maniskill/pn_colorjitter.py"])
def test_widened_augmentation_configs_run_through_the_public_call(rel):
    """pn_shift.py / pn_dropout.py: agent from the config, graph-replayed updates with the
    device-side (Philox) draws, finite logged scalars with the reference's keys."""
    from pointcloud_rl_b200.data import FixedBatchMemory
    from pointcloud_rl_b200.synthetic import synthetic_batch

    B, N, A, S = 4, 200, 5, 13
    obs_shape = {"xyz": [3, N], "rgb": [3, N], "seg": [1, N], "agent": S}
    agent = make_agent(rel, obs_shape, A, hidden=64, batch_size=B).to("cuda")
    mem = FixedBatchMemory(synthetic_batch(0, B, N, A, n_seg=1, state_dim=S))
    for u in range(1, 5):
        out = agent.update_parameters(mem, u)
        assert all(np.isfinite(v) for v in out.values())
    assert {"drq/critic_loss", "drq/actor_loss", "drq/alpha_loss", "drq/entropy"} <= set(out)


@pytest.mark.gpu
def test_reference_written_checkpoint_loads_and_training_continues(tmp_path):
    """a .ckpt written by the REFERENCE's save_checkpoint (checkpoint_utils.py:238-266, parameters
    + the three torch Adam state dicts under `actor_optim` / `critic_optim` / `alpha_optim`) loads into the agent --
    weights, per-tensor Adam moments (cross-checked by parameter NAME against what the reference held) and step
    counts -- and the next update reproduces the reference's own third update.  Then the agent's save_checkpoint
    writes the same layout back."""
    from oracle import pointnet_sac_oracle as O
    from pointcloud_rl_b200.checkpoint import load_checkpoint, save_checkpoint

    g = load_golden("ref_checkpoint_drq_small")
    jg = load_golden("drq_jitter_small")
    m = {k: v.item() for k, v in jg["meta"].items()}
    batch = {k: (dict(v) if isinstance(v, dict) else v) for k, v in jg["batch"].items()}
    obs_shape = {k: (list(v.shape[1:]) if v.ndim > 2 else int(v.shape[1])) for k, v in batch["obs"].items()}
    torch.manual_seed(123)  # different initial weights: everything must come from the file
    agent = make_agent("mfrl/drq/maniskill/pn_jitter.py", obs_shape, m["A"], hidden=64, batch_size=m["B"], precision="fp32",
                       use_cuda_graph=False).to("cuda")
    path = os.path.join(GOLDEN, "ref_checkpoint_drq_small.ckpt")
    ck = load_checkpoint(agent, path, sample=batch)
    assert ck["meta"] == {"updates": 2}
    eng = agent.engine
    for ours, val in g["params"].items():
        assert torch.allclose(eng.p[ours].cpu(), torch.from_numpy(val).reshape(eng.p[ours].shape), atol=0), ours
    mviews, vviews = eng.layout.views(eng.adam_m), eng.layout.views(eng.adam_v)
    steps = eng.steps.cpu().tolist()
    checked = 0
    for ours, ref_name in O.reference_key_map().items():
        for idx, opt in enumerate(("critic_optim", "actor_optim", "alpha_optim")):
            entry = g["adam"][opt].get(ref_name)
            if entry is None:
                continue
            assert torch.equal(mviews[ours].cpu().flatten(), torch.from_numpy(entry["m"]).flatten()), (opt, ours)
            assert torch.equal(vviews[ours].cpu().flatten(), torch.from_numpy(entry["v"]).flatten()), (opt, ours)
            assert steps[idx] == int(entry["step"]), (opt, steps, entry["step"])
            checked += 1
    assert checked == 24 + 6 + 1  # every tensor of the three optimizers (optimizer_utils.py:31-64: one group per tensor)
    # the reference's third update, from the loaded state
    noise = {k: torch.from_numpy(v).cuda() for k, v in g["noise3"].items()}
    eng.upload_batch(batch)
    eng.update(3, noise)
    ret = eng.read_scalars(3)
    ref = {f"{a}/{b}": float(v) for a, sub in g["ret3"].items() for b, v in sub.items()}
    assert set(ret) == set(ref)
    for key, val in ref.items():
        assert ret[key] == pytest.approx(val, rel=1e-3, abs=1e-4), key
    # and back: same keys, optimizer entries in torch.optim.Adam's layout
    out = tmp_path / "ours.ckpt"
    save_checkpoint(agent, out, meta={"updates": 3})
    ours_ck = torch.load(out, map_location="cpu", weights_only=True)
    assert set(ours_ck["state_dict"]) == set(ck["state_dict"])
    for opt in ("actor_optim", "critic_optim", "alpha_optim"):
        a, b = ours_ck["state_dict"][opt], ck["state_dict"][opt]
        assert len(a["param_groups"]) == len(b["param_groups"]) and set(a["state"]) == set(b["state"])
        assert [grp["params"] for grp in a["param_groups"]] == [grp["params"] for grp in b["param_groups"]]
        for i in a["state"]:
            assert a["state"][i]["exp_avg"].shape == b["state"][i]["exp_avg"].shape


@pytest.mark.gpu
def test_rollout_path_bf16_cached_weights_and_fused_inference_aug():
    """SURVEY.md section 8f.4 (BaseAgent.forward, module_utils.py:147-159; DrQ.forward, drq.py:33-44) at the agent's
    precision: bf16 rollout features within 2e-2 of the oracle; the packed weight images are cached between calls and
    rebuilt after an update; inference_aug="same" runs the jitter inside the staging kernel."""
    from oracle import pointnet_sac_oracle as O
    from pointcloud_rl_b200.data import FixedBatchMemory
    from pointcloud_rl_b200.synthetic import synthetic_batch

    N, A, S = 300, 5, 13
    obs_shape = {"xyz": [3, N], "rgb": [3, N], "seg": [1, N], "agent": S}
    agent = make_agent("mfrl/drq/maniskill/pn_jitter.py", obs_shape, A, hidden=64, batch_size=4, precision="bf16",
                       use_cuda_graph=False).to("cuda")
    rs = np.random.RandomState(1)
    obs = O.synthetic_obs(rs, 3, N, n_seg=1, state_dim=S)

    def oracle_mean():
        sd = {k: v.detach().cpu() for k, v in agent.state_dict().items()}
        p = O.params_from_reference_state_dict(sd)
        f = O.pointnet_forward(p, O.preprocess({k: torch.from_numpy(v) for k, v in obs.items() if k != "agent"}))
        return torch.tanh(O.mlp3(p, "actor", torch.cat([f, torch.from_numpy(obs["agent"])], -1))[:, :A])

    pn = agent.actor.backbone.visual_nn
    mean = agent(obs, mode="eval")
    assert float((mean.cpu() - oracle_mean()).abs().max()) < 2e-2
    runner = pn._runner
    n0 = runner.L.launches
    agent(obs, mode="eval")
    per_call_cached = runner.L.launches - n0
    mem = FixedBatchMemory(synthetic_batch(0, 4, N, A, n_seg=1, state_dim=S))
    agent.update_parameters(mem, 1)  # weights move -> the cached images are stale
    n0 = runner.L.launches
    mean2 = agent(obs, mode="eval")
    assert runner.L.launches - n0 == per_call_cached + 1  # exactly one extra C-ABI call: the re-pack
    assert float((mean2.cpu() - oracle_mean()).abs().max()) < 2e-2
    assert not torch.equal(mean, mean2)
    # inference-time augmentation fused into the staging kernel
    aug_agent = make_agent("mfrl/drq/maniskill/pn_jitter.py", obs_shape, A, hidden=64, batch_size=4, precision="bf16",
                           inference_aug="same", use_cuda_graph=False).to("cuda")
    a1, a2 = aug_agent(obs, mode="eval"), aug_agent(obs, mode="eval")
    assert a1.shape == (3, A) and torch.isfinite(a1).all() and not torch.equal(a1, a2)  # fresh jitter every call
    assert float((a1 - a2).abs().max()) < 0.2
    # CUDA-graph replay of the rollout step (the default): same numbers as the eager path, tracks weight updates
    g_agent = make_agent("mfrl/drq/maniskill/pn_jitter.py", obs_shape, A, hidden=64, batch_size=4, precision="bf16").to("cuda")
    g_agent.load_state_dict(agent.state_dict())
    assert torch.equal(g_agent(obs, mode="eval"), agent(obs, mode="eval"))
    assert torch.equal(g_agent(obs, mode="eval"), agent(obs, mode="eval"))  # second call = replay
    g_agent.update_parameters(mem, 1)
    agent2 = make_agent("mfrl/drq/maniskill/pn_jitter.py", obs_shape, A, hidden=64, batch_size=4, precision="bf16",
                        use_cuda_graph=False).to("cuda")
    agent2.load_state_dict(g_agent.state_dict())
    assert torch.equal(g_agent(obs, mode="eval"), agent2(obs, mode="eval"))  # replay sees the updated weights
    s1, s2 = g_agent(obs, mode="explore"), g_agent(obs, mode="explore")
    assert not torch.equal(s1, s2)  # the Philox counter advances inside the graph


def test_lazy_scalars_is_a_dict_that_resolves_on_access_in_update_order():
    """engine.LazyScalars (what update_parameters returns): resolved on first access, earlier updates first."""
    from pointcloud_rl_b200.engine import LazyScalars

    order = []

    class Ev:
        def __init__(self, i):
            self.i = i

        def synchronize(self):
            order.append(self.i)

        def query(self):
            return True

    class Slot:
        def __init__(self, v):
            self.v = np.array([v], np.float32)

        def numpy(self):
            return self.v

    class Eng:
        def __init__(self):
            self._scalar_pending = []

        def _scalars_dict(self, s, updates):
            return {"loss": float(s[0]), "updates": updates}

    eng = Eng()
    rets = [LazyScalars(eng, Slot(float(i)), Ev(i), i) for i in range(4)]
    eng._scalar_pending.extend(rets)
    assert order == []
    assert rets[2]["loss"] == 2.0 and order == [0, 1, 2] and eng._scalar_pending == [rets[3]]
    assert isinstance(rets[3], dict) and dict(rets[3]) == {"loss": 3.0, "updates": 3} and order == [0, 1, 2, 3]
    assert json.loads(json.dumps(rets[0])) == {"loss": 0.0, "updates": 0}
    merged = {}
    merged.update(rets[1])
    assert merged == {"loss": 1.0, "updates": 1} and len(rets[1]) == 2 and "loss" in rets[1]
    assert sorted(rets[0].items()) == [("loss", 0.0), ("updates", 0)] and order == [0, 1, 2, 3]


@pytest.mark.gpu
def test_pipelined_updates_equal_synchronous_updates():
    """update_parameters stages batch i+1 (copy stream + landing buffer) and defers the scalar read-back while update i
    runs: a loop that reads every result late must produce the scalars, weights and alpha of a loop that reads each result
    at once (to run-to-run reproducibility: the gradient-norm and weight-gradient sums are float atomics, so two identical
    synchronous runs already differ in the last bits)."""
    from oracle import pointnet_sac_oracle as O
    from pointcloud_rl_b200.data import DictArray

    B, N, A, S = 6, 96, 5, 13
    obs_shape = {"xyz": [3, N], "rgb": [3, N], "seg": [1, N], "agent": S}
    batches = [O.synthetic_batch(seed=40 + i, B=B, N=N, A=A, n_seg=1, n_pos=0, state_dim=S) for i in range(5)]

    class Memory:
        def __init__(self):
            self.i = 0

        def sample(self, n):
            b = batches[self.i % len(batches)]
            self.i += 1
            return DictArray(b)

    results = {}
    for mode in ("sync", "late"):
        torch.manual_seed(3)
        agent = make_agent("mfrl/drq/maniskill/pn_jitter.py", obs_shape, A, hidden=64, batch_size=B, precision="fp32",
                           seed=11).to("cuda")
        mem, rets = Memory(), []
        for u in range(1, 11):
            r = agent.update_parameters(mem, u)
            if mode == "sync":
                r = dict(r)
            rets.append(r)
        if mode == "late":
            assert len(agent.engine._scalar_pending) == 8  # the ring resolved the two oldest results to reuse their slots
        alpha = agent.alpha
        assert not agent.engine._scalar_pending
        results[mode] = ([dict(r) for r in rets], agent.engine.params.clone(), alpha)
    (rs, ps, a_s), (rl, pl, a_l) = results["sync"], results["late"]
    for u, (a, b) in enumerate(zip(rs, rl)):
        assert a.keys() == b.keys()
        for k in a:
            # last-bit differences grow along the trajectory; an update that saw the wrong batch is off by O(1)
            tol = 1e-3 if u < 6 else 5e-2
            assert abs(a[k] - b[k]) <= tol * max(1.0, abs(a[k])), (u + 1, k, a[k], b[k])
    assert abs(a_s - a_l) <= 1e-4
    assert rs[-1]["drq/alpha"] != rs[0]["drq/alpha"]
