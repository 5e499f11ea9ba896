"""Import the unmodified reference (lz1oceani/pointcloud_rl `pyrl`): from the read-only mount `/root/reference` in the
build container, else from the staged copy `oracle/_ref/` (oracle/build_ref.py; git-ignored, travels to the GPU box).

TEST / BENCH INFRASTRUCTURE ONLY.  Nothing in `pointcloud_rl_b200/` imports this module.
`tests/golden/make_golden.py` uses it to run the reference's own SAC/DrQ/PointNet code on seeded inputs and commits
the results as fixtures that pin `oracle/pointnet_sac_oracle.py`; `bench.py` uses it for the reference arm
(`--impl reference`, `cpu_baseline.kind = "reference"`) and the reference-on-torch-CUDA comparator.

Only stubs for absent third-party packages are provided (oracle/refshim); no reference source is
copied or modified.  Recipe follows SURVEY.md Appendix A.
"""
import os
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
_SHIM = os.path.join(_HERE, "refshim")


def _find_root():
    for cand in (os.environ.get("PCRL_REFERENCE_ROOT"), "/root/reference", os.path.join(_HERE, "_ref")):
        if cand and os.path.isdir(os.path.join(cand, "pyrl")):
            return cand
    return "/root/reference"


REFERENCE_ROOT = _find_root()


def reference_available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "pyrl"))


def load_reference():
    """Returns a namespace of the reference symbols the golden generator needs."""
    if not reference_available():
        raise RuntimeError(f"reference not found at {REFERENCE_ROOT} (nor staged under oracle/_ref: run oracle/build_ref.py "
                           "in the build container)")
    os.environ.setdefault("PYTHONDONTWRITEBYTECODE", "1")
    sys.dont_write_bytecode = True
    for p in (_SHIM, REFERENCE_ROOT):
        if p not in sys.path:
            sys.path.insert(0, p)
    # pyrl/utils/augmentations/image_aug.py:3 imports a module torchvision removed.
    import torchvision.transforms._functional_tensor as _ft

    sys.modules.setdefault("torchvision.transforms.functional_tensor", _ft)

    import pyrl.utils.meta as meta  # noqa: F401
    import pyrl.utils.data as data
    import pyrl.utils.torch  # noqa: F401
    import pyrl.networks as networks
    import pyrl.utils.augmentations  # noqa: F401
    import pyrl.methods.mfrl  # noqa: F401
    from pyrl.methods.builder import build_agent
    from pyrl.networks.utils import get_kwargs_from_shape, replace_placeholder_with_args
    from pyrl.utils.meta import Config
    from gym.spaces import Box

    class NS:
        pass

    ns = NS()
    ns.Config = Config
    ns.build_agent = build_agent
    ns.build_all = networks.build_all
    ns.get_kwargs_from_shape = get_kwargs_from_shape
    ns.replace_placeholder_with_args = replace_placeholder_with_args
    ns.DictArray = data.DictArray
    ns.GDict = data.GDict
    ns.Box = Box
    ns.root = REFERENCE_ROOT
    return ns


def build_reference_agent(ns, cfg_relpath, obs_shape, action_dim, overrides=None):
    """Build the reference agent from one of its own config files (configs/mfrl/...)."""
    import numpy as np

    cfg = ns.Config.fromfile(os.path.join(ns.root, cfg_relpath))
    agent_cfg = cfg.agent_cfg
    for dotted, value in (overrides or {}).items():
        node = agent_cfg
        keys = dotted.split(".")
        for k in keys[:-1]:
            node = node[k]
        node[keys[-1]] = value
    agent_cfg["env_params"] = dict(
        obs_shape=obs_shape,
        action_shape=action_dim,
        action_space=ns.Box(-1.0, 1.0, (action_dim,), dtype=np.float32),
        is_discrete=False,
    )
    cfg = ns.replace_placeholder_with_args(cfg, **ns.get_kwargs_from_shape(obs_shape, action_dim))
    return ns.build_agent(cfg.agent_cfg), cfg
