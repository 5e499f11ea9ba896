"""DrQ + PointNet, colour-jitter augmentation of the point colours, ManiSkill."""
from pointcloud_rl_b200.configs._pn_family import COLOR_JITTER as _CJ, experiment as _experiment

globals().update(_experiment("drq", "maniskill", obs_aug=dict(_CJ), env_name="PushChair_3001-v0"))
