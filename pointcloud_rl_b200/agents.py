"""SAC and DrQ agents with pyrl's public surface (pyrl/methods/mfrl/sac.py:24-214, drq.py:20-165,
pyrl/utils/torch/module_utils.py:112-349), executing the gradient step on the libpcrl kernels.

    agent = build_agent(cfg.agent_cfg)              # MFRL registry, same config dicts
    stats = agent.update_parameters(memory, updates)  # same call, same returned keys
    actions = agent(obs, mode="explore")              # rollout path

Module tree, parameter names and optimizer attribute names match the reference so `state_dict()` keys
line up with its checkpoints.  After construction the parameters are re-pointed into the engine's flat
device buffer: the fused Adam/Polyak kernels and the nn.Module views share storage.
"""
import copy
import re

import numpy as np
import torch
import torch.nn as nn

from .augmentations import build_data_augmentations
from .data import GDict, unwrap
from .engine import AUG_KINDS, HyperParams, PathSpec, UpdateEngine
from .meta import Registry, build_from_cfg
from .networks import ExtendedModule, _mlp_views, _pn_views, build_actor_critic, build_target_network

MFRL = Registry("mfrl")


def build_agent(cfg, default_args=None):
    """pyrl/methods/builder.py:7-11."""
    if cfg["type"] in MFRL:
        return build_from_cfg(cfg, MFRL, default_args)
    return None


class FlatAdam:
    """View of one parameter group of the engine's fused Adam as a torch-style optimizer object: exposes
    `param_groups`, `state_dict()` / `load_state_dict()` in torch.optim.Adam's layout (one group per tensor,
    as build_optimizer does, optimizer_utils.py:31-64) so reference checkpoints round-trip."""

    SUPPORTED = {"type", "lr", "betas", "eps", "weight_decay", "amsgrad", "param_cfg", "constructor"}

    def __init__(self, agent, group, step_index, optim_cfg, names):
        self.agent, self.group, self.step_index, self.names = agent, group, step_index, names
        cfg = dict(optim_cfg)
        unknown = set(cfg) - self.SUPPORTED
        if unknown:
            raise NotImplementedError(f"optimizer options {sorted(unknown)} are not implemented by the fused Adam")
        if cfg.get("type", "Adam") != "Adam" or cfg.get("weight_decay", 0) != 0 or cfg.get("amsgrad", False):
            raise NotImplementedError("the fused optimizer is plain Adam (no weight decay, no amsgrad): "
                                      "optimizer_utils.py:31-64 with the pn_base.py settings")
        self.defaults = dict(lr=float(cfg.get("lr", 1e-3)), betas=tuple(cfg.get("betas", (0.9, 0.999))),
                             eps=float(cfg.get("eps", 1e-8)), weight_decay=0, amsgrad=False)

    @property
    def param_groups(self):
        return [dict(self.defaults, params=[i]) for i in range(len(self.names))]

    def zero_grad(self, set_to_none=False):
        pass

    def state_dict(self):
        eng = self.agent.engine
        state = {}
        if eng is not None:
            m, v = eng.layout.views(eng.adam_m), eng.layout.views(eng.adam_v)
            step = float(eng.steps[self.step_index].item())
            shapes = {k: shp for k, (_, shp) in self.agent._module_params().items()}  # Conv1d weights are [c, C, 1]
            for i, n in enumerate(self.names):
                state[i] = {"step": torch.tensor(step), "exp_avg": m[n].clone().reshape(shapes[n]),
                            "exp_avg_sq": v[n].clone().reshape(shapes[n])}
        return {"state": state, "param_groups": self.param_groups}

    def load_state_dict(self, sd):
        eng = self.agent._ensure_engine()
        for grp in sd.get("param_groups", []):
            # the fused kernel has ONE hyper-parameter set per optimizer: a checkpoint whose groups disagree with the
            # configured values would silently train differently
            for key in ("lr", "betas", "eps"):
                if key in grp and tuple(np.atleast_1d(grp[key]).tolist()) != tuple(np.atleast_1d(self.defaults[key]).tolist()):
                    raise ValueError(f"checkpoint optimizer {key}={grp[key]} differs from the configured {self.defaults[key]}")
            if grp.get("weight_decay", 0) != 0 or grp.get("amsgrad", False):
                raise NotImplementedError("checkpoint optimizer uses weight decay / amsgrad")
        m, v = eng.layout.views(eng.adam_m), eng.layout.views(eng.adam_v)
        steps = set()
        for i, n in enumerate(self.names):
            st = sd["state"].get(i)
            if st is None:
                continue
            m[n].copy_(st["exp_avg"].reshape(m[n].shape))
            v[n].copy_(st["exp_avg_sq"].reshape(v[n].shape))
            steps.add(int(st["step"]))
        if len(steps) > 1:
            raise ValueError("FlatAdam needs one shared step count per optimizer")
        if steps:
            eng.steps[self.step_index] = steps.pop()


class BaseAgent(ExtendedModule):
    def __init__(self):
        super().__init__()
        self._device_ids = None
        self._be_data_parallel = False
        self.obs_processor = None
        self.obs_rms = None
        self.rew_rms = None
        self.batch_size = None
        self.engine = None

    def reset(self, *args, **kwargs):
        pass

    @torch.no_grad()
    def forward(self, obs, **kwargs):
        """Rollout entry (module_utils.py:147-159): obs -> device -> actor(obs, mode=...).  With use_cuda_graph the
        ~15 kernels of one rollout step (stage, fused encode, head, actor MLP, tanh-Gaussian sample) replay from a CUDA
        graph per (batch size, mode): the observation is copied into the graph's static input buffers."""
        kwargs = {k: v for k, v in kwargs.items() if k in ("mode", "num_samples", "aug")}
        with torch.cuda.device(self.device):
            if getattr(self, "use_cuda_graph", False) and self.device.type == "cuda":
                return self._forward_graphed(obs, kwargs)
            obs = GDict(obs).to_torch(device=self.device, non_blocking=True, wrapper=False)
            return self.actor(obs, **kwargs)

    def _forward_graphed(self, obs, kwargs):
        obs = unwrap(obs)
        flat = {k: torch.as_tensor(v) for k, v in obs.items()}
        pn = self.actor.backbone.visual_nn
        # parameter storage is part of the key: the first update re-points the module parameters into the engine's flat
        # buffer (and .to() moves them), which would leave a captured graph reading the old tensors
        key = (tuple((k, tuple(v.shape), v.dtype) for k, v in sorted(flat.items())), kwargs.get("mode", "explore"),
               repr(kwargs.get("aug")), self.actor.backbone.final_mlp.mlp.linear0.weight.data_ptr(),
               pn.conv.mlp.conv0.weight.data_ptr())
        cache = self.__dict__.setdefault("_rollout_graphs", {})
        entry = cache.get(key)
        if entry is None:
            static = {k: torch.empty(v.shape, dtype=v.dtype, device=self.device) for k, v in flat.items()}
            for k, v in flat.items():
                static[k].copy_(v, non_blocking=True)
            side = torch.cuda.Stream(device=self.device)
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):  # warm-up: workspaces, module loading, the packed weight images
                for _ in range(2):
                    self.actor(dict(static), **kwargs)
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize(self.device)
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                out = self.actor(dict(static), **kwargs)
            entry = cache[key] = dict(static=static, graph=graph, out=out, version=pn.weights_version)
        for k, v in flat.items():
            entry["static"][k].copy_(v, non_blocking=True)
        if entry["version"] != pn.weights_version:
            # the packed MMA images of the PointNet weights are graph-external state: refresh them eagerly
            pn.repack()
            entry["version"] = pn.weights_version
        entry["graph"].replay()
        out = entry["out"]
        return [o.clone() for o in out] if isinstance(out, (list, tuple)) else out.clone()

    # data-parallel switches keep their names; the all-reduce lives in the engine (dist.py)
    def to_ddp(self, device_ids=None):
        import torch.distributed as dist

        from .dist import attach

        self._device_ids = device_ids
        self._be_data_parallel = True
        if dist.is_initialized() and self.engine is not None:
            self._attach_ddp(self.engine)
        # otherwise _ensure_engine() attaches as soon as the engine exists (it needs a first batch for its shapes)

    def _attach_ddp(self, eng):
        """What wrapping in DDP does at wrap time (module_utils.py:322-337): every rank starts from rank 0's
        parameters / targets / optimizer state; plus the gradient all-reduce and a per-rank random stream."""
        import torch.distributed as dist

        from .dist import attach, broadcast_state

        if not dist.is_initialized():
            raise RuntimeError("to_ddp() needs torch.distributed to be initialised (one process per GPU)")
        attach(eng)
        broadcast_state(eng)
        eng.seed = int(self.seed) + dist.get_rank()  # augmentation / policy noise must differ between ranks

    def to_normal(self):
        self._be_data_parallel = False

    def recover_ddp(self):
        if self._device_ids is not None:
            self._be_data_parallel = True

    def is_data_parallel(self):
        return self._be_data_parallel

    def no_sync(self, mode="actor"):
        from contextlib import nullcontext

        return nullcontext()


@MFRL.register_module()
class SAC(BaseAgent):
    PREFIX = "sac"

    def __init__(self, actor_cfg, critic_cfg, env_params, batch_size=128, gamma=0.99, reward_scale=1, update_coeff=0.005,
                 alpha=0.2, alpha_optim_cfg=None, automatic_alpha_tuning=True, target_entropy=None, ignore_dones=False,
                 use_episode_dones=False, target_update_interval=1, actor_update_interval=1, shared_backbone=False,
                 shared_target_backbone=None, detach_actor_feature=False, target_smooth=0.90, pre_process=None,
                 precision="bf16", use_cuda_graph=True, seed=0):
        super().__init__()
        if env_params["is_discrete"]:
            raise NotImplementedError("discrete SAC is outside the PointNet continuous-control path")
        if not (shared_backbone and detach_actor_feature) or pre_process is not None:
            raise NotImplementedError("supported: shared_backbone=True, detach_actor_feature=True, pre_process=None "
                                      "(the pn.py / pn_*.py configs)")
        self.is_discrete = False
        self.gamma, self.update_coeff, self.alpha, self.reward_scale = gamma, update_coeff, alpha, reward_scale
        self.ignore_dones, self.batch_size = ignore_dones, batch_size
        self.target_update_interval, self.actor_update_interval = target_update_interval, actor_update_interval
        self.automatic_alpha_tuning, self.shared_backbone = automatic_alpha_tuning, shared_backbone
        self.detach_actor_feature, self.use_episode_dones = detach_actor_feature, use_episode_dones
        self.precision, self.use_cuda_graph, self.seed = precision, use_cuda_graph, seed

        actor_cfg, critic_cfg = copy.deepcopy([actor_cfg, critic_cfg])
        self._actor_optim_cfg, self._critic_optim_cfg = actor_cfg.pop("optim_cfg"), critic_cfg.pop("optim_cfg")
        self._alpha_optim_cfg = alpha_optim_cfg or dict(type="Adam", lr=1e-3)
        pc = self._actor_optim_cfg.get("param_cfg") or {}
        if not any(v is None and re.search(pat, "backbone.visual_nn.x") for pat, v in pc.items()):
            raise NotImplementedError("the actor optimizer must exclude visual_nn (param_cfg={'(.*?)visual_nn(.*?)': None})")
        actor_cfg.update(env_params)
        critic_cfg.update(env_params)
        self.env_params = env_params
        self.actor, self.critic = build_actor_critic(actor_cfg, critic_cfg, shared_backbone)
        self.actor.backbone.visual_nn.precision = precision  # rollout-path encodes use the agent's precision
        shared_target = shared_backbone if shared_target_backbone is None else shared_target_backbone
        if not shared_target:
            raise NotImplementedError("shared_target_backbone=False (a separate, Polyak-averaged target PointNet) is not "
                                      "implemented: the engine encodes next_obs with the live PointNet (builder.py:28-45 "
                                      "with the pn.py configs, where the target shares it)")
        self.target_critic = build_target_network(critic_cfg, self.critic, self.actor, shared_target)

        self.log_alpha = nn.Parameter(torch.ones(1) * float(np.log(np.float32(alpha))))
        self.target_entropy = -float(np.prod(env_params["action_shape"])) if target_entropy is None else target_entropy
        if automatic_alpha_tuning:
            self.alpha = float(self.log_alpha.exp().item())
        tau = update_coeff["default"] if isinstance(update_coeff, dict) else update_coeff
        self._tau = float(tau)  # the visual_nn coefficient is a no-op: the target shares the live PointNet

        names_pn = ["pn.w0", "pn.b0", "pn.w1", "pn.g1", "pn.be1", "pn.w2", "pn.g2", "pn.be2", "pn.wf", "pn.bf", "pn.gf", "pn.bef"]
        mk = ["w0", "b0", "w1", "b1", "w2", "b2"]
        self.actor_optim = FlatAdam(self, "actor", 1, self._actor_optim_cfg, [f"actor.{k}" for k in mk])
        self.critic_optim = FlatAdam(self, "critic", 0, self._critic_optim_cfg,
                                     names_pn + [f"q0.{k}" for k in mk] + [f"q1.{k}" for k in mk])
        self.alpha_optim = FlatAdam(self, "alpha", 2, self._alpha_optim_cfg, ["log_alpha"])
        eps = {o.defaults["eps"] for o in (self.actor_optim, self.critic_optim, self.alpha_optim)}
        if len(eps) != 1:
            raise NotImplementedError("the three optimizers must share one Adam eps")
        self._aug = None
        self._num_aug = 1
        self.actor.backbone.visual_nn.weights_version = 0  # this agent owns the weights and counts their changes

    # ------------------------------------------------------------------ engine plumbing
    def _named_views(self):
        pn = self.actor.backbone.visual_nn
        out = dict(_pn_views(pn))
        out.update(_mlp_views(self.actor.backbone.final_mlp, "actor"))
        for h in (0, 1):
            out.update(_mlp_views(self.critic.values[h].backbone.final_mlp, f"q{h}"))
            out.update(_mlp_views(self.target_critic.values[h].backbone.final_mlp, f"tq{h}"))
        out["log_alpha"] = self.log_alpha.detach()
        return out

    def _hyper(self):
        kind, lo, hi, *rest = (None, 0.0, 0.0) if self._aug is None else self._aug
        axes = rest[0] if (rest and kind == "shift") else 7
        color = tuple(rest[0]) if (rest and kind == "colorjitter") else (0.0, 0.0, 0.0, 0.0)
        head = self.actor.head
        return HyperParams(
            algo=self.PREFIX, gamma=float(self.gamma), reward_scale=float(self.reward_scale), num_aug=self._num_aug,
            aug=kind, aug_lo=lo, aug_hi=hi, aug_axes=axes, aug_color=color, tau=self._tau, actor_update_interval=self.actor_update_interval,
            target_update_interval=self.target_update_interval, lr=self.critic_optim.defaults["lr"],
            actor_lr=self.actor_optim.defaults["lr"], alpha_lr=self.alpha_optim.defaults["lr"],
            betas=self.critic_optim.defaults["betas"], actor_betas=self.actor_optim.defaults["betas"], alpha_betas=self.alpha_optim.defaults["betas"],
            adam_eps=self.critic_optim.defaults["eps"],
            log_std_bound=(head.log_std_min, head.log_std_max), head_scale=head.scale_value, head_bias=head.bias_value,
            target_entropy=float(self.target_entropy), ignore_dones=self.ignore_dones,
            automatic_alpha_tuning=self.automatic_alpha_tuning)

    def _spec(self, sample=None):
        obs_shape = self.env_params["obs_shape"]
        pn = self.actor.backbone.visual_nn
        S = 0
        for key in ("state", "agent"):
            if key in obs_shape:
                S = int(np.prod(obs_shape[key]))
        rgb_u8 = True
        if sample is not None and "rgb" in sample["obs"]:
            rgb_u8 = np.asarray(sample["obs"]["rgb"]).dtype == np.uint8
        mlp = self.actor.backbone.final_mlp.mlp_spec
        return PathSpec(
            n_points=int(obs_shape["xyz"][-1]), action_dim=int(np.prod(self.env_params["action_shape"])), state_dim=S,
            has_rgb="rgb" in obs_shape, rgb_u8=rgb_u8,
            n_pos=int(obs_shape["pos_encoding"][-2]) if "pos_encoding" in obs_shape else 0,
            n_seg=int(obs_shape["seg"][-2]) if "seg" in obs_shape else 0, widths=tuple(pn.mlp_spec),
            out_dim=pn.out_channels, hidden=(mlp[1], mlp[2]), ln_eps=pn.ln_eps)

    def _ensure_engine(self, sample=None):
        if self.engine is not None:
            return self.engine
        device = self.device
        if device.type != "cuda":
            raise RuntimeError("the update path runs on CUDA kernels only: move the agent to a GPU first "
                               "(agent.to('cuda')); there is no CPU fallback")
        eng = UpdateEngine(self._spec(sample), self._hyper(), self.batch_size, device=device, precision=self.precision,
                           seed=self.seed)
        views = self._named_views()
        eng.load_params(views)
        # re-point every module parameter at the engine's flat buffer (shared storage from here on)
        for name, (mod_prm, shape) in self._module_params().items():
            mod_prm.data = eng.p[name].view(shape)
        eng.prime_alpha()
        self.engine = eng
        if self._be_data_parallel:
            self._attach_ddp(eng)
        return eng

    def _module_params(self):
        pn = self.actor.backbone.visual_nn
        m = pn.conv.mlp
        out = {
            "pn.w0": m.conv0.weight, "pn.b0": m.conv0.bias, "pn.w1": m.conv1.weight, "pn.g1": m.norm1.weight,
            "pn.be1": m.norm1.bias, "pn.w2": m.conv2.weight, "pn.g2": m.norm2.weight, "pn.be2": m.norm2.bias,
            "pn.wf": pn.final_mlp[0].weight, "pn.bf": pn.final_mlp[0].bias, "pn.gf": pn.final_mlp[1].weight,
            "pn.bef": pn.final_mlp[1].bias, "log_alpha": self.log_alpha,
        }
        nets = {"actor": self.actor.backbone.final_mlp}
        for h in (0, 1):
            nets[f"q{h}"] = self.critic.values[h].backbone.final_mlp
            nets[f"tq{h}"] = self.target_critic.values[h].backbone.final_mlp
        for net, mlp in nets.items():
            for i in range(3):
                lin = getattr(mlp.mlp, f"linear{i}")
                out[f"{net}.w{i}"] = lin.weight
                out[f"{net}.b{i}"] = lin.bias
        return {k: (v, tuple(v.shape)) for k, v in out.items()}

    def load_state_dict(self, state_dict, strict=True):
        ret = super().load_state_dict(state_dict, strict=strict)
        if self.engine is not None:
            self.engine.refresh_alpha()
            self.engine.prime_alpha()
        self.alpha = float(self.log_alpha.exp().item())
        self._weights_changed()
        return ret

    # ------------------------------------------------------------------ the public step
    def _sample(self, memory):
        batch = memory.sample(self.batch_size)
        if hasattr(batch, "gather_into"):  # device-resident replay ring (replay.py): stays on the GPU
            if self.use_episode_dones:
                raise NotImplementedError("use_episode_dones with the device replay ring")
            return batch
        batch = unwrap(batch)
        if self.use_episode_dones:
            batch["dones"] = batch["episode_dones"]
        return batch

    def update_parameters(self, memory, updates):
        """sac.py:103-214 -- same call, same returned dict; one H2D batch copy in (none with the device replay ring),
        one scalar copy out."""
        batch = self._sample(memory)
        if hasattr(batch, "gather_into"):
            eng = self._ensure_engine(None if self.engine is not None else batch.to_host())
            batch.gather_into(eng)
        else:
            eng = self._ensure_engine(batch)
            eng.upload_batch(batch)
        if self.use_cuda_graph:
            eng.update_graphed(updates)
        else:
            eng.update(updates)
        # the scalars come back asynchronously (engine.LazyScalars: a dict that waits for its device->host copy on first
        # access), so the next call can sample and stage its batch while this update is still running
        ret = eng.read_scalars_async(updates)
        self._weights_changed()
        return ret

    @property
    def alpha(self):
        """exp(log_alpha) as last cached by an update (sac.py:152); waits for updates still in flight."""
        if self.engine is not None:
            self.engine.flush_scalars()
            return self.engine._alpha_before
        return self._alpha

    @alpha.setter
    def alpha(self, v):
        self._alpha = v

    def _weights_changed(self):
        """The rollout path caches the packed MMA images of the PointNet weights: tell it they moved."""
        pn = self.actor.backbone.visual_nn
        pn.weights_version = (pn.weights_version or 0) + 1


@MFRL.register_module()
class DrQ(SAC):
    PREFIX = "drq"

    def __init__(self, num_aug=2, obs_aug=None, svea=False, inference_aug=None, *args, **kwargs):
        super().__init__(*args, **kwargs)
        if svea:
            raise NotImplementedError("SVEA is outside the supported DrQ path (pn_*.py configs use svea=False)")
        self.num_aug, self.svea = num_aug, svea
        self.obs_aug = build_data_augmentations(obs_aug)
        self.inference_aug = self.obs_aug if inference_aug == "same" else build_data_augmentations(inference_aug)
        self._num_aug = int(num_aug)
        if self.obs_aug is not None:
            if len(self.obs_aug) != 1:
                raise NotImplementedError("one fused point-cloud augmentation per agent")
            self._aug = self.obs_aug[0].params()

    @torch.no_grad()
    def forward(self, obs, **kwargs):
        """drq.py:33-44: optional inference-time augmentation, then the SAC rollout path."""
        if self.inference_aug is not None:
            aug = self.inference_aug[0] if len(self.inference_aug) == 1 else None
            if aug is not None and aug.kind in ("jitter", "rot", "shift"):
                # fused into the staging kernel of the encode (pcrl_stage_points), like the update path
                kwargs = dict(kwargs, aug=(AUG_KINDS[aug.kind],) + tuple(aug.params()[1:]))
            else:
                obs = self.inference_aug(GDict(obs).to_torch(device=self.device, wrapper=False))
        return super().forward(obs, **kwargs)
