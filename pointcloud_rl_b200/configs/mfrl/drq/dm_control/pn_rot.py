"""DrQ + PointNet, per-cloud z-rotation augmentation, DM Control."""
from pointcloud_rl_b200.configs._pn_family import ROT_Z as _AUG, experiment as _experiment

globals().update(_experiment("drq", "dm_control", obs_aug=_AUG, env_name="dmc_cheetah_run-v0"))
