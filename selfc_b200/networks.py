"""`define_G(opt)` with the reference's meaning (codes/models/networks.py:12-37), GMM branch only: the options
dict selects `model: SelfC_GMM` -> SelfCInvNet(opt_net, in_nc, out_nc, subnet_type, block_num, down_num)."""
from __future__ import annotations

import math

from .arch import SelfCInvNet

_GMM_MODELS = ("SelfC_GMM",)


def define_G(opt):
    opt_net = opt["network_G"]
    subnet_type = opt_net["which_model_G"]["subnet_type"]
    down_num = int(math.log(opt_net["scale"], 2))
    model_type = opt["model"]
    if model_type not in _GMM_MODELS:
        raise NotImplementedError(
            f"selfc_b200 implements model 'SelfC_GMM' (the SelfC-large rescaler) only, got {model_type!r}; "
            "the Haar 'SelfC'/'IRN' variants and the H.265 codec model are out of scope (SURVEY 8f)")
    return SelfCInvNet(opt_net, opt_net["in_nc"], opt_net["out_nc"], subnet_type, opt_net["block_num"], down_num)
