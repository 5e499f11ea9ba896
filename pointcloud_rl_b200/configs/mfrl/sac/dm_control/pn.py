"""SAC + PointNet on DM Control point clouds (BASELINE config 1)."""
from pointcloud_rl_b200.configs._pn_family import experiment as _experiment

globals().update(_experiment("sac", "dm_control"))
