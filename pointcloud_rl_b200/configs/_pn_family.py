"""One generator for the PointNet SAC/DrQ experiment family (the reference spreads the same values over
configs/mfrl/{sac,drq}/{dm_control,maniskill}/pn*.py and base/pn_base.py).  tests/test_configs.py checks
that every file under configs/mfrl resolves to exactly the dict the reference's file of the same path does."""

SUITES = {
    # per-suite values: PointNet widths / feature size, robot-state placeholder, discount, env + loop settings
    "dm_control": dict(
        widths=[64, 128, 256], feat=50, state="", zero_out=False, critic_bias=True, sac_gamma=0.99, drq_gamma=0.95,
        env=dict(type="gym", env_name="dmc_cheetah_run-v0", obs_mode="pointcloud", stack_frame=3),
        train=dict(total_steps=500000, n_steps=1), drq_train=dict(total_steps=500000, n_steps=1), procs=1, eval_env=None,
    ),
    "maniskill": dict(
        widths=[128, 128, 256], feat=128, state=" + agent_shape", zero_out=True, critic_bias=None, sac_gamma=0.95,
        drq_gamma=0.95,
        env=dict(type="gym", env_name="PushChair_3001-v0", obs_mode="pointcloud", ego_mode=True, no_early_stop=True,
                 with_ext_torque=True, cos_sin_representation=True, reward_scale=0.3),
        train=dict(total_steps=500000, n_steps=4), drq_train=dict(total_steps=1000000, n_steps=4), procs=4,
        eval_env=dict(no_early_stop=False),
    ),
}


def _actor(s):
    mlp = dict(type="LinearMLP", norm_cfg=None, mlp_spec=[f"{s['feat']}{s['state']}" if s["state"] else s["feat"], 1024, 1024,
                                                          "action_shape * 2"], inactivated_output=True)
    if s["zero_out"]:
        mlp["zero_out_indices"] = slice("action_shape", None, None)
    return dict(
        type="ContinuousActor",
        head_cfg=dict(type="TanhGaussianHead", log_std_bound=[-10, 2]),
        nn_cfg=dict(
            type="Visuomotor",
            visual_nn_cfg=dict(type="PointNet", feat_dim="pcd_all_channel", mlp_spec=list(s["widths"]),
                               out_channels=s["feat"], feature_transform=[], ignore_first_ln=True),
            mlp_cfg=mlp,
        ),
        optim_cfg=dict(type="Adam", lr=1e-3, param_cfg={"(.*?)visual_nn(.*?)": None}),
    )


def _critic(s):
    mlp = dict(type="LinearMLP", norm_cfg=None, mlp_spec=[f"{s['feat']}{s['state']} + action_shape", 1024, 1024, 1])
    if s["critic_bias"] is not None:
        mlp["bias"] = s["critic_bias"]
    mlp["inactivated_output"] = True
    return dict(type="ContinuousCritic", num_heads=2,
                nn_cfg=dict(type="Visuomotor", visual_nn_cfg=None, mlp_cfg=mlp), optim_cfg=dict(type="Adam", lr=1e-3))


def experiment(algo, suite, obs_aug=None, env_name=None):
    s = SUITES[suite]
    agent = dict(
        type={"sac": "SAC", "drq": "DrQ"}[algo], batch_size=256, gamma=s[f"{algo}_gamma"], alpha=0.1,
        automatic_alpha_tuning=True, ignore_dones=False, update_coeff={"default": 0.01, "(.*?)visual_nn(.*?)": 0.05},
        target_update_interval=2, actor_update_interval=2, alpha_optim_cfg=dict(type="Adam", lr=1e-3, betas=(0.5, 0.999)),
        shared_backbone=True, detach_actor_feature=True, actor_cfg=_actor(s), critic_cfg=_critic(s),
    )
    env = dict(s["env"])
    if algo == "drq":
        agent.update(num_aug=2, svea=False)
        env.pop("env_name")
        if obs_aug is not None:
            agent["obs_aug"] = obs_aug
    if env_name is not None:
        env["env_name"] = env_name
    loop = s["drq_train"] if algo == "drq" else s["train"]
    evalc = dict(type="Evaluation", num_procs=1, num=1, use_hidden_state=False, save_traj=False, save_video=True,
                 log_every_step=False)
    if s["eval_env"] is not None:
        evalc["env_cfg"] = dict(s["eval_env"])
    return dict(
        agent_cfg=agent,
        env_cfg=env,
        train_cfg=dict(on_policy=False, total_steps=loop["total_steps"], warm_steps=1000, n_steps=loop["n_steps"],
                       n_updates=1, n_eval=-1, n_checkpoint=100000, exp_logger_cfg=dict(type="aim", log_dir="./")),
        replay_cfg=dict(type="ReplayMemory", capacity=100000, sampling_cfg=dict(type="OneStepTransition")),
        rollout_cfg=dict(type="Rollout", num_procs=s["procs"]),
        eval_cfg=evalc,
    )


JITTER = dict(type="RandomJitterPoints", main_key="xyz", req_keys=["xyz"], jitter_range=[-0.01, 0.01])
ROT_Z = dict(type="GlobalRotScaleTrans", main_key="xyz", req_keys=["xyz"], rot_range=[-0.15, 0.15],
             scale_ratio_range=None, translation_range=None, shift_height=False)


def shift(translation_range):
    return dict(type="GlobalRotScaleTrans", main_key="xyz", req_keys=["xyz"], rot_range=None, scale_ratio_range=None,
                translation_range=list(translation_range), shift_height=True)


COLOR_JITTER = dict(type="ColorJitterPoints", main_key="rgb", req_keys=["rgb"], brightness=0.4, contrast=0.4,
                    saturation=0.4, hue=0.5)


def dropout(req_keys):
    return dict(type="RandomDownSample", main_key="xyz", req_keys=list(req_keys), drop_ratio=0.3, fixed_ratio=False)
