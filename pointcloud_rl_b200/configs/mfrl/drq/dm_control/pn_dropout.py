"""DrQ + PointNet, random point down-sampling augmentation, DM Control."""
from pointcloud_rl_b200.configs._pn_family import dropout as _dropout, experiment as _experiment

globals().update(_experiment("drq", "dm_control", obs_aug=_dropout(["xyz", "rgb", "pos_encoding"]), env_name="dmc_cheetah_run-v0"))
