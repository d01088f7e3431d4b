"""Seeded synthetic weights in the reference's state_dict layout (the bundled checkpoint is absent).

U(-b, b) with b = gain / sqrt(fan_in) -- the PyTorch default Conv/Linear init the reference effectively keeps --
drawn with numpy's PCG64 in parameter-registration order, so the same tensors come out on any machine.  Used by bench.py's
product arm; the parity tests draw their weights from the oracle (which implements the same recipe independently).
"""
from __future__ import annotations

import math
import os
from collections import OrderedDict

import numpy as np
import torch


def seeded_state_dict(net: torch.nn.Module, seed: int = 0, gain: float = 1.0) -> "OrderedDict[str, torch.Tensor]":
    rng = np.random.Generator(np.random.PCG64(seed))
    sd: "OrderedDict[str, torch.Tensor]" = OrderedDict()
    fan = {}
    for name, p in net.state_dict().items():
        shape = tuple(p.shape)
        base = name.rsplit(".", 1)[0]
        if name.endswith(".weight"):
            fan[base] = int(np.prod(shape[1:]))
        b = gain / math.sqrt(fan[base])
        sd[name] = torch.from_numpy(rng.uniform(-b, b, size=shape).astype(np.float32))
    return sd


CKPT_ENV = "SELFC_CKPT"
CKPT_DEFAULT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "pretrained_models", "selfc_large_pretrain.pth")


def checkpoint_path():
    """Path of a real SelfC-large checkpoint when one is present on this machine: $SELFC_CKPT, else the reference's bundled
    location pretrained_models/selfc_large_pretrain.pth next to the package (test_SelfC_large_vid4.yml: pretrain_model_G).
    None when there is none (the blob is absent from the reference mount)."""
    for cand in (os.environ.get(CKPT_ENV), CKPT_DEFAULT):
        if cand and os.path.isfile(cand):
            return cand
    return None


def load_checkpoint(path: str) -> "OrderedDict[str, torch.Tensor]":
    """A checkpoint in the reference's format (base_model.py:87-107: a plain state_dict whose keys may carry DataParallel's
    'module.' prefix) -> fp32 CPU tensors under the bare SelfCInvNet names."""
    raw = torch.load(path, map_location="cpu", weights_only=True)
    if isinstance(raw, dict) and "state_dict" in raw and not any(k.endswith(".weight") for k in raw):
        raw = raw["state_dict"]
    sd: "OrderedDict[str, torch.Tensor]" = OrderedDict()
    for k, v in raw.items():
        sd[k[7:] if k.startswith("module.") else k] = v.detach().to(torch.float32).contiguous()
    return sd


def bench_state_dict(net: torch.nn.Module, seed: int = 0):
    """(state_dict, description): the real checkpoint when checkpoint_path() finds one, seeded random weights otherwise."""
    path = checkpoint_path()
    if path is not None:
        sd = load_checkpoint(path)
        net.load_state_dict(sd, strict=True)          # fails loudly on a foreign checkpoint
        return sd, f"checkpoint {path}"
    return seeded_state_dict(net, seed), "seeded random, reference state_dict layout"


def synthetic_net(train: bool = False):
    """SelfCInvNet built from the packaged options file (same keys as the reference's YAML) with seeded weights loaded."""
    import os
    from . import networks, options
    here = os.path.dirname(os.path.abspath(__file__))
    yml = "selfc_large_train_synthetic.yml" if train else "selfc_large_synthetic.yml"
    opt = options.dict_to_nonedict(options.parse(os.path.join(here, "configs", yml), is_train=train))
    return networks.define_G(opt), opt
