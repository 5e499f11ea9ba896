"""DrQ + PointNet, per-cloud translation augmentation (x and z only), DM Control."""
from pointcloud_rl_b200.configs._pn_family import experiment as _experiment, shift as _shift

globals().update(_experiment("drq", "dm_control", obs_aug=_shift([0.04, 0, 0.04]), env_name="dmc_cheetah_run-v0"))
