"""Minimal stand-in for the `addict` package (absent from this image).

TEST INFRASTRUCTURE ONLY: lets oracle/ref_loader.py import the read-only reference
(`pyrl.utils.meta.config` does `from addict import Dict`) so golden vectors can be
generated from the reference's own code.  Not imported by the product package.
"""
import copy


class Dict(dict):
    def __init__(self, *args, **kwargs):
        super().__init__()
        for arg in args:
            if not arg:
                continue
            if isinstance(arg, dict):
                for k, v in arg.items():
                    self[k] = self._hook(v)
            else:
                for k, v in arg:
                    self[k] = self._hook(v)
        for k, v in kwargs.items():
            self[k] = self._hook(v)

    @classmethod
    def _hook(cls, item):
        if isinstance(item, dict) and not isinstance(item, cls):
            return cls(item)
        if isinstance(item, (list, tuple)):
            return type(item)(cls._hook(e) for e in item)
        return item

    def __setattr__(self, name, value):
        self[name] = value

    def __setitem__(self, name, value):
        super().__setitem__(name, value)

    def __getattr__(self, item):
        return self.__getitem__(item)

    def __missing__(self, name):
        value = self.__class__()
        self[name] = value
        return value

    def __delattr__(self, name):
        del self[name]

    def to_dict(self):
        out = {}
        for k, v in self.items():
            if isinstance(v, Dict):
                out[k] = v.to_dict()
            elif isinstance(v, (list, tuple)):
                out[k] = type(v)(e.to_dict() if isinstance(e, Dict) else e for e in v)
            else:
                out[k] = v
        return out

    def copy(self):
        return copy.copy(self)

    def deepcopy(self):
        return copy.deepcopy(self)

    def __copy__(self):
        new = self.__class__()
        for k, v in self.items():
            dict.__setitem__(new, k, v)
        return new

    def __deepcopy__(self, memo):
        new = self.__class__()
        memo[id(self)] = new
        for k, v in self.items():
            dict.__setitem__(new, copy.deepcopy(k, memo), copy.deepcopy(v, memo))
        return new

    def update(self, *args, **kwargs):
        other = {}
        if args:
            other.update(args[0])
        other.update(kwargs)
        for k, v in other.items():
            if k in self and isinstance(self[k], dict) and isinstance(v, dict):
                self[k].update(v)
            else:
                self[k] = self._hook(v)

    def __getstate__(self):
        return dict(self)

    def __setstate__(self, state):
        for k, v in state.items():
            dict.__setitem__(self, k, v)
