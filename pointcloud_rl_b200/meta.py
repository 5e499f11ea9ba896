"""Registry / config plumbing with the same surface as pyrl's (pyrl/utils/meta/registry.py:4-136,
pyrl/utils/meta/config.py, pyrl/networks/utils.py:24-119) so the reference's python-dict configs and
`build_from_cfg(cfg, registry)` calls work unchanged against this package.

Written from the behaviour, not the source: a Registry maps names to classes, `build_from_cfg` pops
"type" and calls the class with the remaining keys, configs are python files whose public top-level
names form the dict, `_base_` files are merged first (child keys win, dicts merge recursively,
`_delete_=True` replaces instead of merging), and string placeholders such as "128 + agent_shape" are
substituted and evaluated once the observation/action shapes are known.
"""
import copy
import inspect
import os
import runpy
from numbers import Number


class Registry:
    def __init__(self, name):
        self._name = name
        self._module_dict = {}

    name = property(lambda self: self._name)
    module_dict = property(lambda self: self._module_dict)

    def __len__(self):
        return len(self._module_dict)

    def __contains__(self, key):
        return key in self._module_dict

    def __repr__(self):
        return f"Registry(name={self._name}, items={sorted(self._module_dict)})"

    def get(self, key):
        return self._module_dict.get(key)

    def register_module(self, name=None, force=False, module=None):
        if not isinstance(force, bool):
            raise TypeError(f"force must be a boolean, but got {type(force)}")
        if name is not None and not isinstance(name, str):
            raise TypeError(f"name must be a str, but got {type(name)}")

        def _register(cls):
            if not (inspect.isclass(cls) or inspect.isfunction(cls)):
                raise TypeError(f"module must be a class or a function, but got {type(cls)}")
            key = name or cls.__name__
            if not force and key in self._module_dict:
                raise KeyError(f"{key} is already registered in {self._name}")
            self._module_dict[key] = cls
            return cls

        return _register(module) if module is not None else _register


def build_from_cfg(cfg, registry, default_args=None):
    if cfg is None:
        return None
    if not isinstance(cfg, dict):
        raise TypeError(f"cfg must be a dict, but got {type(cfg)}")
    if not isinstance(registry, Registry):
        raise TypeError(f"registry must be a Registry, but got {type(registry)}")
    if default_args is not None and not isinstance(default_args, dict):
        raise TypeError(f"default_args must be a dict or None, but got {type(default_args)}")
    args = dict(cfg)
    for k, v in (default_args or {}).items():
        args.setdefault(k, v)
    if "type" not in args:
        raise KeyError(f'`cfg` or `default_args` must contain the key "type", but got {cfg}\n{default_args}')
    obj_type = args.pop("type")
    if isinstance(obj_type, str):
        cls = registry.get(obj_type)
        if cls is None:
            raise KeyError(f"{obj_type} is not in the {registry.name} registry")
    elif inspect.isclass(obj_type):
        cls = obj_type
    else:
        raise TypeError(f"type must be a str or valid type, but got {type(obj_type)}")
    return cls(**args)


class ConfigDict(dict):
    """dict with attribute access; missing keys raise (KeyError / AttributeError) instead of auto-creating."""

    def __init__(self, *args, **kwargs):
        super().__init__()
        for k, v in dict(*args, **kwargs).items():
            self[k] = v

    @classmethod
    def _wrap(cls, v):
        if isinstance(v, dict) and not isinstance(v, ConfigDict):
            return cls(v)
        if isinstance(v, (list, tuple)):
            return type(v)(cls._wrap(e) for e in v)
        return v

    def __setitem__(self, k, v):
        super().__setitem__(k, self._wrap(v))

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError:
            raise AttributeError(f"'{type(self).__name__}' object has no attribute '{k}'") from None

    def __setattr__(self, k, v):
        self[k] = v

    def __delattr__(self, k):
        del self[k]

    def update(self, *args, **kwargs):
        for k, v in dict(*args, **kwargs).items():
            self[k] = v

    def copy(self):
        return type(self)(self)

    def __deepcopy__(self, memo):
        out = type(self)()
        memo[id(self)] = out
        for k, v in self.items():
            dict.__setitem__(out, copy.deepcopy(k, memo), copy.deepcopy(v, memo))
        return out

    def to_dict(self):
        def plain(v):
            if isinstance(v, dict):
                return {k: plain(x) for k, x in v.items()}
            if isinstance(v, (list, tuple)):
                return type(v)(plain(e) for e in v)
            return v

        return plain(self)


BASE_KEY, DELETE_KEY = "_base_", "_delete_"


def merge_dicts(child, base):
    """child over base: dict values merge recursively unless the child dict carries _delete_=True."""
    out = copy.deepcopy(base)
    for k, v in child.items():
        if isinstance(v, dict) and isinstance(out.get(k), dict) and not v.get(DELETE_KEY, False):
            out[k] = merge_dicts(v, out[k])
        else:
            v = copy.deepcopy(v)
            if isinstance(v, dict):
                v.pop(DELETE_KEY, None)
            out[k] = v
    return out


class Config:
    """Python-file configs with `_base_` inheritance and dotted-key overrides."""

    def __init__(self, cfg_dict=None, filename=None):
        object.__setattr__(self, "_cfg", ConfigDict(cfg_dict or {}))
        object.__setattr__(self, "filename", filename)

    @staticmethod
    def _load(filename):
        filename = os.path.abspath(os.path.expanduser(filename))
        if not os.path.isfile(filename):
            raise FileNotFoundError(filename)
        ns = runpy.run_path(filename)
        cfg = {k: v for k, v in ns.items()
               if (k == BASE_KEY or not k.startswith("_")) and not inspect.ismodule(v) and not inspect.isfunction(v)
               and not inspect.isclass(v)}
        bases = cfg.pop(BASE_KEY, [])
        if isinstance(bases, str):
            bases = [bases]
        merged = {}
        for b in bases:
            merged = merge_dicts(Config._load(os.path.join(os.path.dirname(filename), b)), merged)
        return merge_dicts(cfg, merged)

    @classmethod
    def fromfile(cls, filename):
        return cls(cls._load(filename), filename=filename)

    def merge_from_dict(self, options):
        """{"agent_cfg.batch_size": 64} style overrides (run_rl.py --cfg-options)."""
        for dotted, value in options.items():
            node = self._cfg
            keys = dotted.split(".")
            for k in keys[:-1]:
                node = node.setdefault(k, ConfigDict())
            node[keys[-1]] = value

    def dict(self):
        return self._cfg

    def __getattr__(self, k):
        return getattr(self._cfg, k)

    def __getitem__(self, k):
        return self._cfg[k]

    def __setitem__(self, k, v):
        self._cfg[k] = v

    def __setattr__(self, k, v):
        self._cfg[k] = v

    def __contains__(self, k):
        return k in self._cfg

    def __repr__(self):
        return f"Config(path={self.filename}): {self._cfg!r}"


# ----------------------------------------------------------------------------------------------
# shape placeholders (pyrl/networks/utils.py:24-119)
# ----------------------------------------------------------------------------------------------


def get_kwargs_from_shape(obs_shape, action_shape):
    """{"action_shape", "agent_shape", "pcd_all_channel", ...} from env obs/action shapes; channel counts of
    point-cloud leaves are shape[-2] (channel-major [C, N] leaves)."""
    kw = {}
    if action_shape is not None:
        kw["action_shape"] = copy.deepcopy(action_shape)
    if not isinstance(obs_shape, dict):
        kw["obs_shape"] = copy.deepcopy(obs_shape)
        return kw
    for key in ("state", "agent"):
        if key in obs_shape:
            kw["agent_shape"] = obs_shape[key]
    if "xyz" in obs_shape:
        xyz_rgb = sum(obs_shape[k][-2] for k in ("xyz", "rgb") if k in obs_shape)
        total = xyz_rgb + sum(obs_shape[k][-2] for k in ("pos_encoding", "seg") if k in obs_shape)
        if "seg" in obs_shape:
            kw["num_objs"] = obs_shape["seg"][-2]
        kw.update(pcd_all_channel=total, pcd_xyz_rgb_channel=xyz_rgb, pcd_xyz_channel=3)
    return kw


def replace_placeholder_with_args(parameters, **kwargs):
    """Recursively substitute names such as "action_shape" inside strings and evaluate the result
    ("128 + agent_shape" -> 234); slices are rebuilt from their substituted bounds."""
    if parameters is None or isinstance(parameters, Number):
        return parameters
    if isinstance(parameters, Config):
        for k, v in list(parameters.dict().items()):
            parameters[k] = replace_placeholder_with_args(v, **kwargs)
        return parameters
    if isinstance(parameters, dict):
        for k, v in list(parameters.items()):
            parameters[k] = replace_placeholder_with_args(v, **kwargs)
        return parameters
    if isinstance(parameters, (list, tuple)):
        return type(parameters)(replace_placeholder_with_args(v, **kwargs) for v in parameters)
    if isinstance(parameters, slice):
        return slice(*(replace_placeholder_with_args(v, **kwargs) for v in (parameters.start, parameters.stop, parameters.step)))
    if isinstance(parameters, str):
        text = parameters
        for key, val in kwargs.items():
            if key in text:
                text = text.replace(key, str(val))
        try:
            value = eval(text, {"__builtins__": {}}, {})
        except Exception:
            return text
        return text if callable(value) else value
    return parameters
