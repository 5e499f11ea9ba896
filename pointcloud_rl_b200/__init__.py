"""pointcloud_rl_b200: B200-native (sm_100a) PointNet SAC/DrQ update path behind pyrl's registry/config API.

    from pointcloud_rl_b200 import Config, build_agent, replace_placeholder_with_args, get_kwargs_from_shape
    cfg = Config.fromfile(config_path("mfrl/drq/maniskill/pn_jitter.py"))
    ...
    agent = build_agent(cfg.agent_cfg).to("cuda")
    stats = agent.update_parameters(replay, updates)

Importing the package does not need a GPU; running it does (there is no CPU fallback)."""
import os

from .meta import Config, ConfigDict, Registry, build_from_cfg, get_kwargs_from_shape, replace_placeholder_with_args  # noqa: F401

_LAZY = {
    "NETWORK": "networks", "REGRESSION": "networks", "APPLICATION": "networks", "build_all": "networks",
    "PointNet": "networks", "MFRL": "agents", "build_agent": "agents", "SAC": "agents", "DrQ": "agents",
    "AUGMENTATIONS": "augmentations", "build_data_augmentations": "augmentations", "UpdateEngine": "engine",
    "PathSpec": "engine", "HyperParams": "engine", "GDict": "data", "DictArray": "data",
}


def __getattr__(name):
    if name in _LAZY:
        import importlib

        return getattr(importlib.import_module(f".{_LAZY[name]}", __name__), name)
    raise AttributeError(name)


def config_path(rel):
    return os.path.join(os.path.dirname(os.path.abspath(__file__)), "configs", rel)
