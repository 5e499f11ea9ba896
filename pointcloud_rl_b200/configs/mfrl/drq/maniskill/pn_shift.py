"""DrQ + PointNet, per-cloud translation augmentation, ManiSkill."""
from pointcloud_rl_b200.configs._pn_family import experiment as _experiment, shift as _shift

globals().update(_experiment("drq", "maniskill", obs_aug=_shift([0.1, 0.1, 0.1]), env_name="OpenCabinetDrawer_1000-v0"))
