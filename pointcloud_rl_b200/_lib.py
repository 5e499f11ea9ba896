"""ctypes binding of libpcrl.so.  Prototypes are parsed from include/pcrl.h so the header is the single
source of truth.  There is NO fallback: if the library is missing or a symbol is absent this raises."""
import ctypes
import os
import re

import torch

PKG = os.path.dirname(os.path.abspath(__file__))
HEADER = os.path.join(os.path.dirname(PKG), "include", "pcrl.h")
LIB_PATH = os.path.join(PKG, "libpcrl.so")

_CT = {
    "int": ctypes.c_int, "int64_t": ctypes.c_int64, "uint64_t": ctypes.c_uint64, "uint32_t": ctypes.c_uint32,
    "int32_t": ctypes.c_int32, "float": ctypes.c_float,
}


def parse_header(path=HEADER):
    """-> {name: (restype, [(ctype, argname), ...])} for every `pcrl_*` prototype in the header."""
    text = open(path).read()
    text = re.sub(r"/\*.*?\*/", " ", text, flags=re.S)
    text = re.sub(r"//[^\n]*", " ", text)
    protos = {}
    for m in re.finditer(r"(const\s+char\s*\*|int64_t|int)\s+(pcrl_\w+)\s*\(([^;{]*?)\)\s*;", text, flags=re.S):
        ret, name, args = m.group(1), m.group(2), m.group(3).strip()
        restype = ctypes.c_char_p if "char" in ret else _CT[ret]
        argl = []
        if args and args != "void":
            for a in args.split(","):
                a = " ".join(a.split())
                if "*" in a:
                    argl.append((ctypes.c_void_p, a.split("*")[-1].strip()))
                else:
                    ty, an = a.replace("const ", "").rsplit(" ", 1)
                    argl.append((_CT[ty.strip()], an))
        protos[name] = (restype, argl)
    return protos


_SYNC_DEBUG = bool(os.environ.get("PCRL_SYNC"))


class PcrlError(RuntimeError):
    pass


class _Lib:
    def __init__(self):
        if not os.path.exists(LIB_PATH):
            raise PcrlError(
                f"{LIB_PATH} not found: build it with `python -m pointcloud_rl_b200.build` "
                "(there is no CPU or PyTorch fallback for this path)"
            )
        self.cdll = ctypes.CDLL(LIB_PATH)
        self.protos = parse_header()
        for name, (restype, args) in self.protos.items():
            fn = getattr(self.cdll, name)  # AttributeError if the .so does not export a declared symbol
            fn.restype = restype
            fn.argtypes = [t for t, _ in args]
        if self.cdll.pcrl_abi_version() != 1:
            raise PcrlError("libpcrl.so ABI version mismatch; rebuild")
        self.launches = 0  # number of C-ABI calls that launch kernels (bench's gpu_launches claim)

    def last_error(self):
        return self.cdll.pcrl_last_error().decode()

    def __getattr__(self, name):
        full = "pcrl_" + name
        protos = self.__dict__["protos"]
        if full not in protos:
            raise AttributeError(name)
        fn = getattr(self.cdll, full)
        restype, args = protos[full]
        is_status = restype is ctypes.c_int and name not in ("abi_version", "sm_count", "set_strict_tf32", "destroy",
                                                             "host_memcpy_mt")

        def call(*a):
            conv = []
            for v, (t, _) in zip(a, args):
                if t is ctypes.c_void_p:
                    if v is None:
                        conv.append(None)
                    elif torch.is_tensor(v):
                        conv.append(v.data_ptr())
                    else:
                        conv.append(int(v))
                else:
                    conv.append(v)
            if len(conv) != len(args):
                raise TypeError(f"{full} takes {len(args)} arguments, got {len(conv)}")
            rc = fn(*conv)
            if _SYNC_DEBUG and is_status:  # PCRL_SYNC=1: find the launch an asynchronous CUDA error belongs to
                try:
                    torch.cuda.synchronize()
                except Exception as e:  # noqa: BLE001
                    raise PcrlError(f"{full}{tuple(x if not isinstance(x, int) or x < 1 << 32 else hex(x) for x in conv)}: {e}") from e
            if is_status:
                self.launches += 1
                if rc != 0:
                    raise PcrlError(f"{full} failed (code {rc}): {self.last_error()}")
            return rc

        self.__dict__[name] = call
        return call


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = _Lib()
    return _lib


def stream_ptr(stream=None):
    s = stream if stream is not None else torch.cuda.current_stream()
    return s.cuda_stream
