"""Seeded synthetic weights in the reference's state_dict layout (the bundled checkpoint is absent).

U(-b, b) with b = gain / sqrt(fan_in) -- the PyTorch default Conv/Linear init the reference effectively keeps --
drawn with numpy's PCG64 in parameter-registration order, so the same tensors come out on any machine.  Used by bench.py's
product arm; the parity tests draw their weights from the oracle (which implements the same recipe independently).
"""
from __future__ import annotations

import math
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


def synthetic_net(train: bool = False):
    """SelfCInvNet built from the packaged options file (same keys as the reference's YAML) with seeded weights loaded."""
    import os
    from . import networks, options
    here = os.path.dirname(os.path.abspath(__file__))
    yml = "selfc_large_train_synthetic.yml" if train else "selfc_large_synthetic.yml"
    opt = options.dict_to_nonedict(options.parse(os.path.join(here, "configs", yml), is_train=train))
    return networks.define_G(opt), opt
