"""Point-cloud augmentation registry entries (pyrl/utils/augmentations/{builder,pcd_aug}.py).

On the update path these objects are *descriptors*: DrQ reads them and the staging kernel applies the
augmentation fused into the load (pcrl_stage_points).  Calling one directly on a dict of device tensors
(DrQ.inference_aug, drq.py:33-44) runs the same kernel-side arithmetic via torch ops on xyz only."""
import torch

from .meta import Registry, build_from_cfg

AUGMENTATIONS = Registry("data augmentation")


class BaseAugmentation:
    kind = None  # engine aug name

    def __init__(self, main_key=None, req_keys=None):
        self.main_key = main_key
        self.req_keys = req_keys or [main_key]
        assert main_key in self.req_keys, f"{main_key}, {req_keys} do not satisfy the requirement!"

    def params(self):
        raise NotImplementedError


@AUGMENTATIONS.register_module()
class RandomJitterPoints(BaseAugmentation):
    """xyz += U(lo, hi) i.i.d. per coordinate per point (pcd_aug.py:307-322)."""

    kind = "jitter"

    def __init__(self, main_key="inputs/xyz", req_keys=None, jitter_range=[-0.1, 0.1]):
        super().__init__(main_key, req_keys)
        self.jitter_range = [float(jitter_range[0]), float(jitter_range[1])]

    def params(self):
        return self.kind, self.jitter_range[0], self.jitter_range[1]

    def __call__(self, data):
        data = dict(data)
        for key in self.req_keys:
            if key in data:
                x = data[key]
                lo, hi = self.jitter_range
                data[key] = x + (torch.rand_like(x) * (hi - lo) + lo)
        return data

    def __repr__(self):
        return f"RandomJitterPoints(jitter_range={self.jitter_range})"


@AUGMENTATIONS.register_module()
class GlobalRotScaleTrans(BaseAugmentation):
    """Per-cloud rigid augmentation (pcd_aug.py:126-215).  Supported forms: rotation only about z by U(rot_range)
    (`pn_rot.py`) and translation only by U(-t, t) per axis with `shift_height=True` (`pn_shift.py`; the enabled axes
    must share one magnitude, e.g. [0.1, 0.1, 0.1] or [0.04, 0, 0.04]).  Scaling, rotation+translation and the
    `shift_height=False` quirk (the reference zeroes the LAST CLOUD's translation, pcd_aug.py:195-196) raise."""

    def __init__(self, main_key=None, req_keys=None, rot_range=[-0.78539816, 0.78539816], rot_axis="z",
                 scale_ratio_range=[0.95, 1.05], translation_range=[0, 0, 0], shift_height=False):
        main_key = main_key[0] if isinstance(main_key, (list, tuple)) else main_key
        super().__init__(main_key, req_keys)
        if scale_ratio_range is not None:
            raise NotImplementedError("GlobalRotScaleTrans: scaling is not supported")
        if (rot_range is None) == (translation_range is None):
            raise NotImplementedError("GlobalRotScaleTrans: exactly one of rot_range / translation_range (pn_rot.py, pn_shift.py)")
        if rot_range is not None:
            self.kind = "rot"
            if rot_axis not in ("z", 2):
                raise NotImplementedError("GlobalRotScaleTrans: only rot_axis='z'")
            if not isinstance(rot_range, (list, tuple)):
                rot_range = [-rot_range, rot_range]
            self.rot_range = [float(rot_range[0]), float(rot_range[1])]
        else:
            self.kind = "shift"
            if not shift_height:
                raise NotImplementedError("GlobalRotScaleTrans: shift_height=False (zeroes the last cloud's shift) is not supported")
            t = [float(v) for v in translation_range]
            mags = sorted({v for v in t if v != 0.0})
            if len(t) != 3 or len(mags) != 1 or mags[0] < 0:
                raise NotImplementedError("GlobalRotScaleTrans: the shifted axes must share one positive range")
            self.shift = mags[0]
            self.axes = sum(1 << i for i, v in enumerate(t) if v != 0.0)

    def params(self):
        if self.kind == "rot":
            return self.kind, self.rot_range[0], self.rot_range[1]
        return self.kind, -self.shift, self.shift, self.axes

    def __call__(self, data):
        data = dict(data)
        draw = None
        for key in self.req_keys:
            if key in data:
                x = data[key]
                if self.kind == "rot":
                    if draw is None:
                        draw = torch.empty(x.shape[0], device=x.device).uniform_(*self.rot_range)
                    c, s = torch.cos(draw)[:, None], torch.sin(draw)[:, None]
                    data[key] = torch.stack([c * x[:, 0] - s * x[:, 1], s * x[:, 0] + c * x[:, 1], x[:, 2]], dim=1)
                else:
                    if draw is None:
                        draw = (torch.rand(x.shape[0], 3, device=x.device) - 0.5) * 2 * self.shift
                        draw = draw * torch.tensor([float(bool(self.axes & (1 << i))) for i in range(3)], device=x.device)
                    data[key] = x + draw[:, :, None]
        return data

    def __repr__(self):
        if self.kind == "rot":
            return f"GlobalRotScaleTrans(rot_range={self.rot_range})"
        return f"GlobalRotScaleTrans(translation={self.shift}, axes={self.axes:03b})"


@AUGMENTATIONS.register_module()
class RandomDownSample(BaseAugmentation):
    """One random subset of the points per call, shared by all clouds and applied to every point-cloud key
    (pcd_aug.py:228-268; `pn_dropout.py`: drop_ratio=0.3, fixed_ratio=False).  The update path keeps the staged shape:
    dropped points are replaced by a kept one, which leaves the max-pooled features unchanged."""

    kind = "downsample"

    def __init__(self, main_key="inputs/xyz", req_keys=None, max_num_points=None, drop_ratio=None, fixed_ratio=True):
        super().__init__(main_key, req_keys)
        if max_num_points is not None or drop_ratio is None:
            raise NotImplementedError("RandomDownSample: only the drop_ratio form (pn_dropout.py) is supported")
        if not 0.0 <= float(drop_ratio) < 1.0:
            raise ValueError("drop_ratio must be in [0, 1)")
        self.drop_ratio, self.fixed_ratio = float(drop_ratio), bool(fixed_ratio)

    def params(self):
        return self.kind, self.drop_ratio, float(self.fixed_ratio)

    def __call__(self, data):
        data = dict(data)
        index = None
        for key in self.req_keys:
            if key in data:
                x = data[key]
                if index is None:
                    N = x.shape[-1]
                    hi = int(N * self.drop_ratio)
                    n_drop = hi if self.fixed_ratio else (int(torch.randint(hi, (1,)).item()) if hi > 0 else 0)
                    index = torch.rand(N, device=x.device).argsort()[: N - n_drop]
                data[key] = x[..., index]
        return data

    def __repr__(self):
        return f"RandomDownSample(drop_ratio={self.drop_ratio}, fixed_ratio={self.fixed_ratio})"


@AUGMENTATIONS.register_module()
class ColorJitterPoints(BaseAugmentation):
    """torchvision ColorJitter on the uint8 point colours viewed as a [B', 3, 1, N] image batch (pcd_aug.py:269-303;
    `pn_colorjitter.py`: brightness = contrast = saturation = 0.4, hue = 0.5).  One random op order and one factor
    per op are drawn per CALL and shared by the whole batch (torchvision semantics).  Runs as pcrl_color_jitter_points."""

    kind = "colorjitter"

    def __init__(self, main_key="inputs/rgb", req_keys="inputs/rgb", brightness=0.5, contrast=0.5, saturation=0.5, hue=0.5):
        req_keys = [req_keys] if isinstance(req_keys, str) else req_keys
        super().__init__(main_key, req_keys)
        if brightness < 0 or brightness > 1:
            raise ValueError("brightness shoud be non-negative")
        if contrast < 0 or contrast > 1:
            raise ValueError("contrast shoud be non-negative")
        if saturation < 0 or saturation > 1:
            raise ValueError("saturation shoud be non-negative")
        if hue < 0 or hue > 0.5:
            raise ValueError("hue shoud be non-negative")
        self.brightness, self.contrast, self.saturation, self.hue = float(brightness), float(contrast), float(saturation), float(hue)

    def params(self):
        return self.kind, 0.0, 0.0, (self.brightness, self.contrast, self.saturation, self.hue)

    def __call__(self, data):
        """Rollout-side use (DrQ.inference_aug): the same kernel on device uint8 colours."""
        from ._lib import lib, stream_ptr

        data = dict(data)
        for key in self.req_keys:
            if key in data:
                x = data[key]
                if x.dtype != torch.uint8 or not x.is_cuda:
                    raise NotImplementedError("ColorJitterPoints runs on device uint8 colours [B,3,N]")
                x = x.contiguous()
                out = torch.empty_like(x)
                if not hasattr(self, "_counter"):
                    self._counter = torch.zeros(1, dtype=torch.int64, device=x.device)
                with torch.cuda.device(x.device):
                    lib().color_jitter_points(x, x.shape[0], x.shape[-1], None, self.brightness, self.contrast,
                                              self.saturation, self.hue, 0x5EED, self._counter, 6, out, stream_ptr())
                self._counter.add_(1)
                data[key] = out
        return data

    def __repr__(self):
        return (f"ColorJitterPoints(brightness={self.brightness},contrast={self.contrast},saturation={self.saturation},"
                f"hue={self.hue})")


class DataAugmentations:
    def __init__(self, transforms):
        self.transforms = [build_from_cfg(t, AUGMENTATIONS) if isinstance(t, dict) else t for t in transforms]

    def __call__(self, data):
        for t in self.transforms:
            data = t(data)
        return data

    def __getitem__(self, i):
        return self.transforms[i]

    def __len__(self):
        return len(self.transforms)


def build_data_augmentations(cfg, default_args=None):
    if cfg is None:
        return None
    return DataAugmentations(cfg if isinstance(cfg, (list, tuple)) else [cfg])
