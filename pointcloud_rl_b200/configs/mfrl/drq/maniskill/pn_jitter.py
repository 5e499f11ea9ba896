"""DrQ + PointNet, per-point jitter augmentation, ManiSkill (BASELINE config 2)."""
from pointcloud_rl_b200.configs._pn_family import JITTER as _AUG, experiment as _experiment

globals().update(_experiment("drq", "maniskill", obs_aug=_AUG, env_name="OpenCabinetDrawer_1000-v0"))
