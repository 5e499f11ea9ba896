"""SAC + PointNet on ManiSkill point clouds."""
from pointcloud_rl_b200.configs._pn_family import experiment as _experiment

globals().update(_experiment("sac", "maniskill"))
