"""Host-side mirror checks (CPU): registry / config semantics, and that this package's config files resolve
to exactly what the reference's files of the same path resolve to (container only: needs /root/reference)."""
import os

import pytest

from pointcloud_rl_b200 import Config, config_path, get_kwargs_from_shape, replace_placeholder_with_args
from pointcloud_rl_b200.meta import ConfigDict, Registry, build_from_cfg, merge_dicts

FILES = ["mfrl/sac/dm_control/pn.py", "mfrl/sac/maniskill/pn.py", "mfrl/drq/maniskill/pn_jitter.py",
         "mfrl/drq/maniskill/pn_rot.py", "mfrl/drq/dm_control/pn_jitter.py", "mfrl/drq/dm_control/pn_rot.py",
         "mfrl/drq/maniskill/pn_shift.py", "mfrl/drq/dm_control/pn_shift.py",
         "mfrl/drq/maniskill/pn_dropout.py", "mfrl/drq/dm_control/pn_dropout.py",
         "mfrl/drq/maniskill/pn_colorjitter.py", "mfrl/drq/dm_control/pn_colorjitter.py"]


def _plain(x):
    if isinstance(x, dict):
        return {k: _plain(v) for k, v in x.items()}
    if isinstance(x, (list, tuple)):
        return [_plain(v) for v in x]
    if isinstance(x, slice):
        return ("slice", x.start, x.stop, x.step)
    return x


@pytest.mark.parametrize("rel", FILES)
def test_config_matches_reference_file(rel):
    from oracle.ref_loader import load_reference, reference_available

    if not reference_available():
        pytest.skip("reference tree not present on this machine")
    ns = load_reference()
    ref = ns.Config.fromfile(os.path.join(ns.root, "configs", rel))
    ours = Config.fromfile(config_path(rel))
    for key in ("agent_cfg", "env_cfg", "train_cfg", "replay_cfg", "rollout_cfg", "eval_cfg"):
        assert _plain(ours[key]) == _plain(ref[key].to_dict() if hasattr(ref[key], "to_dict") else ref[key]), key


def test_placeholders_resolve_like_the_reference():
    cfg = Config.fromfile(config_path("mfrl/drq/maniskill/pn_jitter.py"))
    obs_shape = {"xyz": [3, 1200], "rgb": [3, 1200], "seg": [1, 1200], "agent": 106}
    kw = get_kwargs_from_shape(obs_shape, 22)
    assert kw == {"action_shape": 22, "agent_shape": 106, "num_objs": 1, "pcd_all_channel": 7, "pcd_xyz_rgb_channel": 6,
                  "pcd_xyz_channel": 3}
    cfg = replace_placeholder_with_args(cfg, **kw)
    a = cfg.agent_cfg
    assert a.actor_cfg.nn_cfg.visual_nn_cfg.feat_dim == 7
    assert a.actor_cfg.nn_cfg.mlp_cfg.mlp_spec == [234, 1024, 1024, 44]
    assert a.actor_cfg.nn_cfg.mlp_cfg.zero_out_indices == slice(22, None, None)
    assert a.critic_cfg.nn_cfg.mlp_cfg.mlp_spec == [256, 1024, 1024, 1]


def test_registry_and_build_from_cfg():
    reg = Registry("things")

    @reg.register_module()
    class A:
        def __init__(self, x, y=2):
            self.x, self.y = x, y

    with pytest.raises(KeyError):
        reg.register_module(module=A)
    reg.register_module(name="B", module=A)
    obj = build_from_cfg(dict(type="A", x=1), reg, default_args=dict(y=5))
    assert (obj.x, obj.y) == (1, 5)
    assert build_from_cfg(None, reg) is None
    with pytest.raises(KeyError):
        build_from_cfg(dict(type="nope"), reg)
    with pytest.raises(TypeError):
        build_from_cfg([1], reg)


def test_base_merge_and_delete():
    base = dict(a=dict(x=1, y=2), b=3)
    assert merge_dicts(dict(a=dict(y=5)), base) == dict(a=dict(x=1, y=5), b=3)
    assert merge_dicts(dict(a=dict(_delete_=True, z=9)), base) == dict(a=dict(z=9), b=3)
    c = ConfigDict(a=dict(b=1))
    assert c.a.b == 1
    with pytest.raises(AttributeError):
        c.missing
