"""DrQ + PointNet, per-cloud z-rotation augmentation, ManiSkill."""
from pointcloud_rl_b200.configs._pn_family import ROT_Z as _AUG, experiment as _experiment

globals().update(_experiment("drq", "maniskill", obs_aug=_AUG))
