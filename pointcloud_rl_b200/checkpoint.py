"""Checkpoint I/O in the reference's layout (pyrl/utils/torch/checkpoint_utils.py:148-266).

A reference `.ckpt` is `torch.save({"meta": ..., "state_dict": sd})` where `sd` holds every parameter under the module
tree's names (`actor.backbone.visual_nn.conv.mlp.conv0.weight`, ...; the shared PointNet appears under the actor, both
critic heads and both target heads) AND, under the optimizer attribute names (`actor_optim`, `critic_optim`,
`alpha_optim`), each optimizer's `torch.optim.Adam.state_dict()` (one param group per tensor, in
`named_parameters()` order: optimizer_utils.py:31-64).  These functions read and write exactly that, mapping the
per-tensor Adam state to and from the engine's flat `m` / `v` buffers (agents.FlatAdam).

    save_checkpoint(agent, "model.ckpt")            # loadable by the reference's load_checkpoint
    load_checkpoint(agent, "model_100000.ckpt")     # a file the reference's save_checkpoint wrote
"""
import os
from collections import OrderedDict

import torch

OPTIMIZERS = ("actor_optim", "critic_optim", "alpha_optim")


def get_state_dict(agent):
    """checkpoint_utils.get_state_dict: parameters + the optimizers' state dicts under their attribute names."""
    sd = OrderedDict((k, v.detach().cpu()) for k, v in agent.state_dict().items())
    for name in OPTIMIZERS:
        opt = getattr(agent, name, None)
        if opt is not None:
            osd = opt.state_dict()
            for st in osd["state"].values():
                for k, v in st.items():
                    if torch.is_tensor(v):
                        st[k] = v.detach().cpu()
            sd[name] = osd
    return sd


def save_checkpoint(agent, filename, meta=None):
    """checkpoint_utils.save_checkpoint (:238-266)."""
    if meta is not None and not isinstance(meta, dict):
        raise TypeError(f"meta must be a dict or None, but got {type(meta)}")
    d = os.path.dirname(str(filename))
    if d:
        os.makedirs(d, exist_ok=True)
    with open(filename, "wb") as f:
        torch.save({"meta": meta or {}, "state_dict": get_state_dict(agent)}, f)
        f.flush()


def load_state_dict(agent, state_dict, strict=False, sample=None):
    """checkpoint_utils.load_state_dict (:23-93): module parameters through nn.Module.load_state_dict, optimizer entries
    through the optimizers' own load_state_dict.  `sample` (a replay batch) is only needed when the agent has not seen a
    batch yet and the observation dtype cannot be told from env_params (float colours)."""
    state_dict = OrderedDict(state_dict)
    optim = {name: state_dict.pop(name) for name in OPTIMIZERS if name in state_dict}
    if list(state_dict.keys())[0].startswith("module."):
        state_dict = OrderedDict((k[7:], v) for k, v in state_dict.items())
    ret = agent.load_state_dict(state_dict, strict=strict)
    if optim:
        agent._ensure_engine(sample)  # the Adam moments live in the engine's flat buffers
        for name, osd in optim.items():
            getattr(agent, name).load_state_dict(osd)
    return ret


def load_checkpoint(agent, filename, map_location="cpu", strict=False, sample=None):
    """checkpoint_utils.load_checkpoint (:148-178) for a local file; returns the checkpoint dict."""
    if not os.path.isfile(str(filename)):
        raise IOError(f"{filename} is not a checkpoint file")
    checkpoint = torch.load(str(filename), map_location=map_location, weights_only=False)
    if not isinstance(checkpoint, dict):
        raise RuntimeError(f"No state_dict found in checkpoint file {filename}")
    load_state_dict(agent, checkpoint.get("state_dict", checkpoint), strict=strict, sample=sample)
    return checkpoint
