"""Stub of `torchviz` (import-only dependency of pyrl.utils.torch.comp_graph). Test infrastructure only."""


def make_dot(*args, **kwargs):
    raise RuntimeError("torchviz is stubbed out in the oracle shim")
