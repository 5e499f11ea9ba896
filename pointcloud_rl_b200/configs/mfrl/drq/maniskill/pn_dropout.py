"""DrQ + PointNet, random point down-sampling augmentation, ManiSkill."""
from pointcloud_rl_b200.configs._pn_family import dropout as _dropout, experiment as _experiment

globals().update(_experiment("drq", "maniskill", obs_aug=_dropout(["xyz", "rgb", "seg"]), env_name="OpenCabinetDrawer_1000-v0"))
