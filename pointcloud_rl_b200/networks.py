"""Host-side mirror of the reference's network registry for the PointNet SAC/DrQ path.

Same registry names, constructor kwargs and `state_dict` keys as pyrl:
  NETWORK["PointNet" | "LinearMLP" | "ConvMLP" | "Visuomotor"], REGRESSION["TanhGaussianHead"],
  APPLICATION["ContinuousActor" | "ContinuousCritic"], build_all(cfg).
The modules are parameter containers with the reference's module tree (so checkpoints line up:
`actor.backbone.visual_nn.conv.mlp.conv0.weight`, `critic.values.0.backbone.final_mlp.mlp.linear1.bias`, ...);
their forward passes run the libpcrl kernels (no autograd -- gradients exist only inside
`agent.update_parameters`, computed by the hand-written backward kernels).  Options outside the supported
subset raise instead of silently differing.
"""
import copy
from typing import Dict

import numpy as np
import torch
import torch.nn as nn

from ._lib import lib, stream_ptr
from .engine import PathSpec
from .meta import ConfigDict, Registry, build_from_cfg

NETWORK = Registry("neural_network")
REGRESSION = Registry("regression")
APPLICATION = Registry("application")


def build_all(cfg, default_args=None):
    """pyrl/networks/builder.py:11-22."""
    if cfg is None:
        return None
    if isinstance(cfg, (list, tuple)):
        return [build_all(c, default_args) for c in cfg]
    for reg in (NETWORK, REGRESSION, APPLICATION):
        if cfg["type"] in reg.module_dict:
            return build_from_cfg(cfg, reg, default_args)
    raise RuntimeError(f"No this model type:{cfg['type']}!")


class ExtendedModule(nn.Module):
    """The bits of pyrl's ExtendedModule (module_utils.py:11-68) callers rely on."""

    is_recurrent = False

    def __init__(self):
        super().__init__()
        self._in_test = False

    def set_mode(self, mode="train"):
        self._in_test = mode == "test"
        for m in self.children():
            if isinstance(m, ExtendedModule):
                m.set_mode(mode)
        return self

    @property
    def device(self):
        return next(self.parameters()).device

    @property
    def dtype(self):
        return next(self.parameters()).dtype

    @property
    def trainable_parameters(self):
        return [p for p in self.parameters() if p.requires_grad]

    @property
    def num_trainable_parameters(self):
        return sum(p.numel() for p in self.parameters() if p.requires_grad)

    @property
    def size_trainable_parameters(self):
        return sum(p.numel() * p.element_size() for p in self.parameters() if p.requires_grad)

    @property
    def grad_norm(self):
        gs = [p.grad.detach().norm(2) for p in self.parameters() if p.requires_grad and p.grad is not None]
        return float(torch.stack(gs).norm(2)) if gs else 0.0

    def pop_attr(self, name):
        ret = getattr(self, name, None)
        if hasattr(self, name):
            setattr(self, name, None)
        return ret

    def no_sync(self):
        from contextlib import nullcontext

        return nullcontext()


def _check(cond, msg):
    if not cond:
        raise NotImplementedError(f"pointcloud_rl_b200 supports the PointNet SAC/DrQ path only: {msg}")


def _is_ln1d(norm_cfg):
    return norm_cfg is not None and norm_cfg.get("type") in ("LN1d", "LNkd", "LayerNorm1D")


class _MLPBase(ExtendedModule):
    def __init__(self, mlp_spec, block, norm_cfg, act_cfg, bias, inactivated_output, ignore_first_ln, zero_out_indices,
                 **kwargs):
        super().__init__()
        _check(not kwargs or set(kwargs) <= {"dense_init_cfg", "separate_module", "nn_cfg"}, f"MLP options {sorted(kwargs)}")
        _check(kwargs.get("dense_init_cfg") is None, "dense_init_cfg")
        _check(act_cfg is None or act_cfg.get("type") == "ReLU", f"activation {act_cfg}")
        self.mlp_spec = [int(x) for x in mlp_spec]
        self.mlp = nn.Sequential()
        n = len(self.mlp_spec) - 1
        for i in range(n):
            last = inactivated_output and i == n - 1
            use_norm = norm_cfg is not None and not last and not (ignore_first_ln and i == 0)
            use_bias = (not use_norm) if bias == "auto" else bool(bias)
            if block == "Conv":
                self.mlp.add_module(f"conv{i}", nn.Conv1d(self.mlp_spec[i], self.mlp_spec[i + 1], 1, bias=use_bias))
            else:
                self.mlp.add_module(f"linear{i}", nn.Linear(self.mlp_spec[i], self.mlp_spec[i + 1], bias=use_bias))
            if use_norm:
                self.mlp.add_module(f"norm{i}", nn.LayerNorm(self.mlp_spec[i + 1], eps=norm_cfg.get("eps", 1e-5)))
            if not last and act_cfg is not None:
                self.mlp.add_module(f"act{i}", nn.ReLU(inplace=True))
        if zero_out_indices is not None:  # mlp.py:78-83: near-zero log-std rows of the actor's last layer
            last_dense = [m for m in self.mlp if isinstance(m, (nn.Linear, nn.Conv1d))][-1]
            with torch.no_grad():
                last_dense.weight[zero_out_indices].uniform_(-1e-3, 1e-3)
                last_dense.bias[zero_out_indices].uniform_(-1e-3, 1e-3)


@NETWORK.register_module()
class LinearMLP(_MLPBase):
    """Linear-ReLU-...-Linear (mlp.py:98-100).  Supported: norm_cfg=None, ReLU, bias, 2 hidden layers."""

    def __init__(self, mlp_spec, norm_cfg=None, act_cfg=dict(type="ReLU"), bias="auto", inactivated_output=True,
                 zero_out_indices=None, ignore_first_ln=False, **kwargs):
        _check(norm_cfg is None, "LinearMLP with a norm layer")
        _check(len(mlp_spec) == 4 and inactivated_output, "LinearMLP must be [in, h1, h2, out] with inactivated_output")
        super().__init__(mlp_spec, "Linear", None, act_cfg, bias, inactivated_output, False, zero_out_indices, **kwargs)


@NETWORK.register_module()
class ConvMLP(_MLPBase):
    """Conv1d(k=1)-LN1d-ReLU blocks (mlp.py:104-108).  Supported: 3 layers, LN1d, ignore_first_ln, ReLU."""

    def __init__(self, mlp_spec, norm_cfg=dict(type="LN1d"), act_cfg=dict(type="ReLU"), bias="auto",
                 inactivated_output=True, ignore_first_ln=False, zero_out_indices=None, **kwargs):
        _check(_is_ln1d(norm_cfg), f"per-point norm {norm_cfg} (only LN1d)")
        _check(len(mlp_spec) == 4 and ignore_first_ln and not inactivated_output,
               "ConvMLP must be [C, c1, c2, c3] with ignore_first_ln=True, inactivated_output=False")
        super().__init__(mlp_spec, "Conv", norm_cfg, act_cfg, bias, inactivated_output, ignore_first_ln, zero_out_indices,
                         **kwargs)


class KernelRunner:
    """Inference-time launcher shared by the module forwards: stages observations, runs the PointNet
    encode, the 3-layer MLPs and the tanh-Gaussian head through the C ABI.  Workspaces are cached per
    batch size.  (Rollout path: BaseAgent.forward -> actor(obs), module_utils.py:147-159.)"""

    def __init__(self, precision="bf16", seed=0):
        self.precision = precision
        self.tf32 = 1 if precision in ("bf16", "tf32") else 0
        self.seed = seed
        self.L = lib()
        self._ws = {}
        self._counter = None
        self.calls = 0

    def _workspace(self, spec: PathSpec, B, device):
        key = (B, spec.n_points, spec.C, spec.widths, spec.out_dim, str(device))
        ws = self._ws.get(key)
        if ws is None:
            L = lib()
            c1, c2, c3 = spec.widths
            f32 = dict(dtype=torch.float32, device=device)
            ws = {
                "xf": torch.zeros(B, spec.NP, spec.CP, **f32),
                "pooled": torch.zeros(B, c3, **f32),
                "z": torch.zeros(B, spec.out_dim, **f32),
            }
            if self.precision == "bf16":
                ws["xh"] = torch.zeros(B * spec.NP * 16, dtype=torch.bfloat16, device=device)
                ws["wpack"] = torch.zeros(int(L.pointnet_wpack_bytes(c1, c2, c3)), dtype=torch.uint8, device=device)
                ws["keys"] = torch.zeros(B * c3, dtype=torch.int64, device=device)
            else:
                chunk = max(1, min(B, 32 if self.precision == "tf32" else 64))
                query = L.pointnet_fwd_tf32_workspace if self.precision == "tf32" else L.pointnet_fwd_f32_workspace
                ws["fwd_bytes"] = int(query(chunk, spec.NP, c1, c2, c3))
                ws["scratch"] = torch.zeros(ws["fwd_bytes"], dtype=torch.uint8, device=device)
            self._ws[key] = ws
        return ws

    def encode(self, spec: PathSpec, p: Dict[str, torch.Tensor], obs: Dict[str, torch.Tensor], out=None, aug=None,
               weights_version=None):
        """obs: device tensors xyz [B,3,N] f32, rgb [B,3,N] u8|f32, pos_encoding [B,F,N] u8, seg [B,K,N] bool/u8.
        weights_version: an integer the owner bumps whenever the PointNet weights change; the packed MMA images are then
        rebuilt only on a change (None = unknown: rebuilt on every call)."""
        L, st = lib(), stream_ptr()
        xyz = obs["xyz"].contiguous().float()
        B, device = xyz.shape[0], xyz.device
        ws = self._workspace(spec, B, device)
        c1, c2, c3 = spec.widths
        rgb = obs.get("rgb")
        rgb_u8 = rgb is not None and rgb.dtype == torch.uint8
        if rgb is not None:
            rgb = rgb.contiguous() if rgb_u8 else rgb.contiguous().float()
        pos = obs.get("pos_encoding")
        seg = obs.get("seg")
        as_u8 = lambda t: None if t is None else t.contiguous().to(torch.uint8)
        kind, lo, hi, *rest = (0, 0.0, 0.0) if aug is None else aug
        if rest and kind == 3:
            kind |= (int(rest[0]) & 7) << 8
        if self._counter is None:
            self._counter = torch.zeros(1, dtype=torch.int64, device=device)
        L.stage_points(xyz, rgb, int(rgb_u8), as_u8(pos), 0 if pos is None else pos.shape[1], as_u8(seg),
                       0 if seg is None else seg.shape[1], B, spec.n_points, 1, kind, lo, hi, None, self.seed,
                       self._counter, 7, ws["xf"], ws.get("xh"), spec.CP, st)
        if kind:
            self._counter.add_(1)
        if self.precision == "bf16":
            tag = (weights_version, int(rgb_u8))
            if weights_version is None or ws.get("packed") != tag:
                L.pointnet_pack_weights_part(p["pn.w0"], p["pn.b0"], p["pn.w1"], p["pn.g1"], p["pn.be1"], p["pn.w2"],
                                             p["pn.g2"], p["pn.be2"], spec.C, c1, c2, c3, int(rgb_u8), 1, ws["wpack"], st)
                ws["packed"] = tag
            L.pointnet_fwd_bf16(ws["xh"], B, spec.n_points, spec.NP, ws["wpack"], c1, c2, c3, spec.ln_eps, ws["keys"],
                                ws["pooled"], None, st)
        else:
            fwd = L.pointnet_fwd_tf32 if self.precision == "tf32" else L.pointnet_fwd_f32
            fwd(ws["xf"], B, spec.n_points, spec.NP, spec.CP, spec.C, p["pn.w0"], p["pn.b0"], p["pn.w1"], p["pn.g1"],
                p["pn.be1"], p["pn.w2"], p["pn.g2"], p["pn.be2"], c1, c2, c3, spec.ln_eps, ws["pooled"], None,
                ws["scratch"], ws["fwd_bytes"], st)
        D = spec.out_dim
        if out is None:
            out = ws.get("out")
            if out is None:
                out = ws["out"] = torch.empty(B, D, dtype=torch.float32, device=device)
        L.linear_fwd(ws["pooled"], c3, p["pn.wf"], p["pn.bf"], ws["z"], D, B, c3, D, 0, self.tf32, st)
        L.layernorm_fwd(ws["z"], p["pn.gf"], p["pn.bef"], out, out.stride(0), None, None, B, D, spec.head_ln_eps, st)
        self.calls += 1
        return out

    def repack(self, p, weights_version):
        L, st = lib(), stream_ptr()
        for key, ws in self._ws.items():
            if isinstance(ws, dict) and "wpack" in ws and "packed" in ws:
                spec_c, (c1, c2, c3) = key[2], key[3]
                rgb_u8 = ws["packed"][1]
                L.pointnet_pack_weights_part(p["pn.w0"], p["pn.b0"], p["pn.w1"], p["pn.g1"], p["pn.be1"], p["pn.w2"],
                                             p["pn.g2"], p["pn.be2"], spec_c, c1, c2, c3, rgb_u8, 1, ws["wpack"], st)
                ws["packed"] = (weights_version, rgb_u8)

    def mlp3(self, p, net, x, K, nout):
        L, st = lib(), stream_ptr()
        M, dev = x.shape[0], x.device
        h1n, h2n = p[f"{net}.w0"].shape[0], p[f"{net}.w1"].shape[0]
        key = ("mlp", M, h1n, h2n, nout, str(dev))
        bufs = self._ws.get(key)
        if bufs is None:  # activations of the rollout MLP: allocated once per batch size
            bufs = self._ws[key] = tuple(torch.empty(M, n, dtype=torch.float32, device=dev) for n in (h1n, h2n, nout))
        h1, h2, out = bufs
        L.linear_fwd(x, x.stride(0), p[f"{net}.w0"], p[f"{net}.b0"], h1, h1n, M, K, h1n, 1, self.tf32, st)
        L.linear_fwd(h1, h1n, p[f"{net}.w1"], p[f"{net}.b1"], h2, h2n, M, h1n, h2n, 1, self.tf32, st)
        L.linear_fwd(h2, h2n, p[f"{net}.w2"], p[f"{net}.b2"], out, nout, M, h2n, nout, 0, self.tf32, st)
        return out


def _pn_views(pointnet) -> Dict[str, torch.Tensor]:
    m = pointnet.conv.mlp
    w = lambda t: t.detach().reshape(t.shape[0], t.shape[1])
    return {
        "pn.w0": w(m.conv0.weight), "pn.b0": m.conv0.bias.detach(), "pn.w1": w(m.conv1.weight),
        "pn.g1": m.norm1.weight.detach(), "pn.be1": m.norm1.bias.detach(), "pn.w2": w(m.conv2.weight),
        "pn.g2": m.norm2.weight.detach(), "pn.be2": m.norm2.bias.detach(), "pn.wf": pointnet.final_mlp[0].weight.detach(),
        "pn.bf": pointnet.final_mlp[0].bias.detach(), "pn.gf": pointnet.final_mlp[1].weight.detach(),
        "pn.bef": pointnet.final_mlp[1].bias.detach(),
    }


def _mlp_views(mlp, net) -> Dict[str, torch.Tensor]:
    m = mlp.mlp
    return {f"{net}.w0": m.linear0.weight.detach(), f"{net}.b0": m.linear0.bias.detach(),
            f"{net}.w1": m.linear1.weight.detach(), f"{net}.b1": m.linear1.bias.detach(),
            f"{net}.w2": m.linear2.weight.detach(), f"{net}.b2": m.linear2.bias.detach()}


@NETWORK.register_module()
class PointNet(ExtendedModule):
    """pyrl/networks/backbones/pointnet.py:76-157.  Supported subset (everything the pn_*.py configs use):
    feature_transform=[], global_feat=True, LN1d(eps) + ReLU, ignore_first_ln=True, out_channels set."""

    def __init__(self, feat_dim, mlp_spec=[64, 128, 1024], out_channels=None, global_feat=True, feature_transform=[1],
                 norm_cfg=dict(type="LN1d", eps=1e-6), act_cfg=dict(type="ReLU"), ignore_first_ln=False, num_patch=1,
                 precision="bf16", **kwargs):
        super().__init__()
        _check(len(feature_transform) == 0, "PointNet feature_transform (STN branches)")
        _check(global_feat and out_channels is not None, "PointNet needs global_feat=True and out_channels")
        _check(not kwargs, f"PointNet options {sorted(kwargs)}")
        self.feat_dim, self.mlp_spec, self.out_channels = int(feat_dim), [int(c) for c in mlp_spec], int(out_channels)
        self.global_feat, self.feature_transform = global_feat, feature_transform
        self.ln_eps = float(norm_cfg.get("eps", 1e-5))
        self.conv = ConvMLP([self.feat_dim] + self.mlp_spec, norm_cfg=norm_cfg, act_cfg=act_cfg, inactivated_output=False,
                            ignore_first_ln=ignore_first_ln)
        self.final_mlp = nn.Sequential(nn.Linear(self.mlp_spec[-1], self.out_channels), nn.LayerNorm(self.out_channels))
        self.precision = precision
        self._runner = None
        self.weights_version = None  # the owning agent counts weight changes here (None: unknown, re-pack every call)

    def repack(self):
        """Rebuild the cached MMA weight images of every workspace now (used before replaying a captured rollout graph,
        where the per-call version check is not part of the graph)."""
        if self._runner is not None:
            self._runner.repack(_pn_views(self), self.weights_version)

    def spec_for(self, n_points, n_pos=0, n_seg=0, has_rgb=True, rgb_u8=True, **extra):
        return PathSpec(n_points=n_points, action_dim=extra.get("action_dim", 1), state_dim=extra.get("state_dim", 0),
                        has_rgb=has_rgb, rgb_u8=rgb_u8, n_pos=n_pos, n_seg=n_seg, widths=tuple(self.mlp_spec),
                        out_dim=self.out_channels, hidden=extra.get("hidden", (1024, 1024)), ln_eps=self.ln_eps)

    @torch.no_grad()
    def forward(self, inputs, object_feature=True, concat_state=None, aug=None, **kwargs):
        _check(not kwargs, f"PointNet.forward options {sorted(kwargs)}")
        if not isinstance(inputs, dict):  # bare [B,C,N] tensor: xyz + extra float channels
            x = inputs.to(self.device, torch.float32)
            inputs = {"xyz": x[:, :3]}
            if x.shape[1] > 3:
                inputs["rgb"] = x[:, 3:6]
                _check(x.shape[1] <= 6, "bare tensor inputs with more than 6 channels (pass a dict)")
        obs = {k: torch.as_tensor(v).to(self.device) for k, v in inputs.items() if k in ("xyz", "rgb", "pos_encoding", "seg")}
        rgb = obs.get("rgb")
        spec = self.spec_for(obs["xyz"].shape[-1], 0 if "pos_encoding" not in obs else obs["pos_encoding"].shape[1],
                             0 if "seg" not in obs else obs["seg"].shape[1], rgb is not None,
                             rgb is not None and rgb.dtype == torch.uint8)
        if spec.C != self.feat_dim:
            raise ValueError(f"observation has {spec.C} channels, PointNet was built with feat_dim={self.feat_dim}")
        if self._runner is None or self._runner.precision != self.precision:
            self._runner = KernelRunner(self.precision)
        return self._runner.encode(spec, _pn_views(self), obs, aug=aug, weights_version=self.weights_version).clone()


@NETWORK.register_module()
class Visuomotor(ExtendedModule):
    """visual_nn(obs) | robot state | action -> final_mlp  (visuomotor.py:16-146); non-recurrent subset."""

    def __init__(self, visual_nn_cfg, mlp_cfg, rnn_cfg=None, obs_feat_cfg=None, ac_feat_cfg=None, prev_ac_feat_cfg=None,
                 freeze_visual_nn=False, freeze_mlp=False, **kwargs):
        super().__init__()
        _check(rnn_cfg is None and obs_feat_cfg is None and ac_feat_cfg is None and prev_ac_feat_cfg is None,
               "Visuomotor rnn / obs_feat / ac_feat branches")
        _check(not freeze_visual_nn and not freeze_mlp, "frozen sub-networks")
        self.visual_nn = kwargs.get("visual_nn", None) or build_all(visual_nn_cfg)  # may be shared (builder.py:60-66)
        self.final_mlp = build_all(mlp_cfg)
        self.saved_feature = None
        self.saved_visual_feature = None

    @torch.no_grad()
    def forward(self, obs, actions=None, feature=None, visual_feature=None, save_feature=False, detach_visual=False,
                with_robot_state=True, aug=None, **kwargs):
        assert isinstance(obs, dict), f"obs is not a dict! {type(obs)}"
        obs = dict(obs)
        robot_state = None
        for key in ("state", "agent"):
            if key in obs:
                assert robot_state is None, "Please provide only one robot state!"
                robot_state = torch.as_tensor(obs.pop(key)).to(self.device, torch.float32)
        if feature is None:
            feat = self.visual_nn(obs, aug=aug) if visual_feature is None else visual_feature
            if save_feature or visual_feature is not None:
                self.saved_visual_feature = feat.clone()
            if robot_state is not None and with_robot_state:
                feat = torch.cat([feat, robot_state], dim=-1)
            if save_feature:
                self.saved_feature = feat.clone()
        else:
            feat = feature
        if actions is not None:
            feat = torch.cat([feat, torch.as_tensor(actions).to(self.device, torch.float32)], dim=-1)
        feat = feat.contiguous()
        runner = self.visual_nn._runner or KernelRunner(self.visual_nn.precision)
        p = _mlp_views(self.final_mlp, "m")
        return runner.mlp3(p, "m", feat, feat.shape[1], self.final_mlp.mlp_spec[-1]).clone()


@REGRESSION.register_module()
class TanhGaussianHead(ExtendedModule):
    """tanh(Normal(mean, std)) policy head (regression_heads/gaussian.py:70-87); scalar scale/bias bounds."""

    def __init__(self, bound=None, dim_output=None, nn_cfg=None, predict_std=True, init_log_std=-0.5, clip_return=False,
                 num_heads=1, log_std_bound=[-20, 2], epsilon=1e-6):
        super().__init__()
        _check(nn_cfg is None and predict_std and num_heads == 1, "TanhGaussianHead nn_cfg / fixed std / mixtures")
        _check(abs(epsilon - 1e-6) < 1e-12, "TanhGaussianHead epsilon != 1e-6")
        if bound is not None:
            lo, hi = [np.asarray(b, dtype=np.float32).reshape(-1) for b in bound]
            dim_output = lo.shape[0] if dim_output is None else dim_output
            lo, hi = np.broadcast_to(lo, (dim_output,)), np.broadcast_to(hi, (dim_output,))
            self.lb = nn.Parameter(torch.tensor(lo.copy()), requires_grad=False)
            self.ub = nn.Parameter(torch.tensor(hi.copy()), requires_grad=False)
            self.scale = nn.Parameter(torch.tensor((hi - lo) / 2), requires_grad=False)
            self.bias = nn.Parameter(torch.tensor((hi + lo) / 2), requires_grad=False)
            _check(float(np.ptp((hi - lo) / 2)) == 0 and float(np.ptp((hi + lo) / 2)) == 0,
                   "per-dimension action bounds (all dimensions must share one [low, high])")
            self.scale_value, self.bias_value = float(self.scale[0]), float(self.bias[0])
        else:
            self.scale_value, self.bias_value = 1.0, 0.0
        self.dim_output = dim_output
        self.log_std_min, self.log_std_max = float(log_std_bound[0]), float(log_std_bound[1])
        self._counter = None

    @torch.no_grad()
    def forward(self, feature, num_samples=1, mode="explore", **kwargs):
        _check(num_samples == 1, "num_samples > 1")
        A = feature.shape[-1] // 2
        M = feature.shape[0]
        if mode in ("mean", "eval"):
            return torch.tanh(feature[:, :A]) * self.scale_value + self.bias_value
        _check(mode in ("explore", "sample", "max-entropy"), f"head mode {mode}")
        L, st = lib(), stream_ptr()
        dev = feature.device
        if self._counter is None:
            self._counter = torch.zeros(1, dtype=torch.int64, device=dev)
        act = torch.empty(M, A, dtype=torch.float32, device=dev)
        nlp = torch.empty(M, dtype=torch.float32, device=dev)
        eps = torch.empty(M, A, dtype=torch.float32, device=dev)
        L.tanh_gaussian_fwd(feature.contiguous(), M, A, self.log_std_min, self.log_std_max, self.scale_value,
                            self.bias_value, None, 0x5eed, self._counter, 9, act, A, nlp, eps, st)
        self._counter.add_(1)
        return [act, nlp[:, None]] if mode == "max-entropy" else act


class ActorCriticBase(ExtendedModule):
    def __init__(self, nn_cfg=None, head_cfg=None, mlp_cfg=None, backbone=None):
        super().__init__()
        assert nn_cfg is None or backbone is None
        _check(mlp_cfg is None, "ActorCriticBase mlp_cfg")
        self.backbone = build_all(nn_cfg) if backbone is None else backbone
        self.final_mlp = None
        self.head = build_all(head_cfg)

    @torch.no_grad()
    def forward(self, obs, actions=None, **kwargs):
        head_kwargs = {k: kwargs.pop(k) for k in ("mode", "num_samples") if k in kwargs}
        feature = self.backbone(obs, actions, **{k: v for k, v in kwargs.items()
                                                 if k in ("feature", "visual_feature", "save_feature", "detach_visual", "aug")})
        return self.head(feature, **head_kwargs) if self.head is not None else feature


@APPLICATION.register_module(name="ContinuousPolicy")
@APPLICATION.register_module()
class ContinuousActor(ActorCriticBase):
    """applications/actor_critic.py:62-72."""

    def __init__(self, nn_cfg=None, head_cfg=None, mlp_cfg=None, backbone=None, action_space=None, obs_shape=None,
                 action_shape=None, **kwargs):
        head_cfg = copy.deepcopy(head_cfg)
        if head_cfg is not None and action_space is not None and getattr(action_space, "is_bounded", lambda: False)():
            head_cfg["bound"] = [action_space.low, action_space.high]
        super().__init__(nn_cfg=nn_cfg, head_cfg=head_cfg, mlp_cfg=mlp_cfg, backbone=backbone)


@APPLICATION.register_module(name="ContinuousValue")
@APPLICATION.register_module()
class ContinuousCritic(ExtendedModule):
    """num_heads independent value heads, outputs concatenated [B, num_heads] (actor_critic.py:87-133)."""

    def __init__(self, nn_cfg=None, head_cfg=None, mlp_cfg=None, backbone=None, share_feature=False, obs_shape=None,
                 action_shape=None, num_heads=1, average_grad=True, **kwargs):
        super().__init__()
        _check(backbone is None and not share_feature and head_cfg is None and mlp_cfg is None,
               "ContinuousCritic shared-feature / head variants")
        self.num_heads = num_heads
        self.values = nn.ModuleList([ActorCriticBase(nn_cfg=nn_cfg) for _ in range(num_heads)])
        # every head wraps the SAME visual_nn object when one was injected into nn_cfg (builder.py:60-66)

    @torch.no_grad()
    def forward(self, obs, actions=None, **kwargs):
        return torch.cat([v(obs, actions, **kwargs) for v in self.values], dim=-1)


SHARED_KEYS = ["visual_nn"]


def _inject_shared(cfg, source_backbone):
    cfg = copy.deepcopy(cfg)
    nn_cfg = cfg["nn_cfg"]
    for name in SHARED_KEYS:
        item = getattr(source_backbone, name, None)
        if item is not None:
            nn_cfg[f"{name}_cfg"] = None
            dict.__setitem__(nn_cfg, name, item)  # a live module travels inside the cfg (builder.py:60-66)
    return cfg


def build_actor_critic(actor_cfg, critic_cfg, shared_backbone=False):
    """pyrl/networks/builder.py:48-73."""
    actor = build_all(actor_cfg)
    if shared_backbone:
        assert "Visuomotor" in actor_cfg["nn_cfg"]["type"], "Only Visuomotor models can share the visual backbone"
        critic_cfg = _inject_shared(critic_cfg, actor.backbone)
    return actor, build_all(critic_cfg)


def build_target_network(network_cfg, network, shared_network=None, shared_backbone=False):
    """pyrl/networks/builder.py:28-45: target = fresh heads around the live shared PointNet, hard-updated."""
    shared_network = network if shared_network is None else shared_network
    if shared_backbone:
        target = build_all(_inject_shared(network_cfg, shared_network.backbone))
    else:
        target = copy.deepcopy(network)
    src = dict(network.named_parameters())
    shared_ids = {id(p) for p in network.parameters()}
    with torch.no_grad():
        for name, prm in target.named_parameters():
            if id(prm) not in shared_ids:
                prm.copy_(src[name])
                prm.requires_grad_(False)
    return target
