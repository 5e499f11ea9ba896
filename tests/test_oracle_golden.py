"""Pins the CPU oracle (oracle/pointnet_sac_oracle.py) against golden vectors produced by the
reference's own code (tests/golden/make_golden.py).  CPU only."""
import numpy as np
import pytest
import torch

from oracle import pointnet_sac_oracle as O
from tests.conftest import load_golden


def _t(tree):
    return {k: torch.from_numpy(np.asarray(v)) for k, v in tree.items()}


@pytest.mark.parametrize("name", ["pointnet_fwd_c7", "pointnet_fwd_c7_dup", "pointnet_fwd_c9_dmc"])
def test_pointnet_forward_matches_reference(name):
    g = load_golden(name)
    p = _t(g["params"])
    x = O.preprocess(_t(g["obs"]))
    feat, pooled, idx = O.pointnet_forward(p, x, return_pool=True)
    ref_idx = torch.from_numpy(g["idx"])
    ref_pooled = torch.from_numpy(g["pooled"])
    # the restatement sums in a different order than Conv1d, so allow argmax flips only between
    # values that are equal to within a few ulp; everything else must be the same index
    mism = idx != ref_idx
    h = O.pointnet_point_features(p, x)
    if mism.any():
        v_ours = torch.gather(h, 2, idx[..., None])[..., 0][mism]
        v_ref = torch.gather(h, 2, ref_idx[..., None])[..., 0][mism]
        assert torch.allclose(v_ours, v_ref, rtol=1e-5, atol=1e-6)
    assert mism.float().mean() < 0.01
    torch.testing.assert_close(pooled, ref_pooled, rtol=1e-4, atol=1e-5)
    torch.testing.assert_close(feat, torch.from_numpy(g["feat"]), rtol=1e-4, atol=1e-4)


def test_argmax_ties_pick_smallest_index():
    g = load_golden("pointnet_fwd_c7_dup")
    N = g["obs"]["xyz"].shape[-1]
    q = N - N // 4
    # the tail of each cloud duplicates the head, so an argmax in the tail would mean a tie was
    # resolved to the larger index
    assert (g["idx"] < q).all()
    p = _t(g["params"])
    _, _, idx = O.pointnet_forward(p, O.preprocess(_t(g["obs"])), return_pool=True)
    assert (idx < q).all()


@pytest.mark.parametrize("name", ["sac_dmc_small", "drq_jitter_small", "drq_rot_small", "drq_shift_small", "drq_downsample_small",
                                  "drq_colorjitter_small"])
def test_update_matches_reference(name):
    g = load_golden(name)
    m = {k: v.item() for k, v in g["meta"].items()}
    hp = dict(
        algo=m["algo"], gamma=m["gamma"], reward_scale=m["reward_scale"], num_aug=m["num_aug"],
        aug=m["aug"] or None, tau=m["tau"], actor_update_interval=m["actor_update_interval"],
        target_update_interval=m["target_update_interval"], target_entropy=m["target_entropy"],
    )
    state = O.new_state(_t(g["init"]))
    batch = dict(g["batch"])
    for u in range(1, m["n_updates"] + 1):
        ret = O.update(state, batch, u, hp, _t(g[f"noise{u}"]))
        ref = {f"{a}/{b}": v for a, sub in g[f"ret{u}"].items() for b, v in sub.items()}
        assert set(ret.keys()) == set(ref.keys()), (sorted(ret), sorted(ref))
        for key, val in ref.items():
            assert ret[key] == pytest.approx(float(val), rel=2e-4, abs=2e-5), (u, key)
        after = _t(g[f"after{u}"])
        for key, val in after.items():
            # Adam's first steps move every weight by ~lr regardless of gradient scale, so compare
            # the applied delta rather than the raw weight
            torch.testing.assert_close(state["params"][key], val, rtol=1e-4, atol=2e-5, msg=lambda s: f"update {u} {key}: {s}")
