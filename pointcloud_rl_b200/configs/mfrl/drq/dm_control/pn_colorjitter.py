"""DrQ + PointNet, colour-jitter augmentation of the point colours, DM Control."""
from pointcloud_rl_b200.configs._pn_family import COLOR_JITTER as _CJ, experiment as _experiment

globals().update(_experiment("drq", "dm_control", obs_aug=dict(_CJ), env_name="dmc_cheetah_run-v0"))
