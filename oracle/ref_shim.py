"""Import the UNMODIFIED reference from /root/reference.  BUILD-CONTAINER ONLY, TEST INFRASTRUCTURE.

/root/reference does not exist on the GPU box; nothing on the ``-m gpu`` / smoke / bench
path imports this file.  It is used by ``oracle/make_golden.py`` (fixture generation) and
by the optional live cross-check in ``tests/test_oracle_live_reference.py`` (skipped when
the mount is absent).

Two shims, both documented in SURVEY.md 8(c):
  1. stub modules for imageio / lmdb / thop / matplotlib / skvideo (absent here; only
     imported, never used on this path);
  2. ``STPNet.reparametrize`` (models/modules/SelfC_GMM_arch_inv.py:412-417) hard-codes
     ``torch.cuda.FloatTensor`` and draws unseeded noise: replaced by the same formula
     ``eps*exp(logvar)+mu`` with an INJECTED eps of the reference's shape [B,48,5,T,h,w].
"""
from __future__ import annotations

import os
import sys
import types

REF_ROOT = "/root/reference/codes"
VID4_YAML = REF_ROOT + "/options/test/rescaling/test_SelfC_large_vid4.yml"


def available() -> bool:
    return os.path.isdir(REF_ROOT)


def _install_stubs():
    def stub(name, **attrs):
        if name in sys.modules:
            return
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m

    stub("imageio")
    stub("lmdb")
    stub("thop", profile=lambda *a, **k: (0, 0), clever_format=lambda *a, **k: a[0])
    try:
        import matplotlib  # noqa: F401
    except Exception:
        stub("matplotlib")
        stub("matplotlib.pyplot")
        sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
        sys.modules["matplotlib"].use = lambda *a, **k: None
    stub("skvideo")
    stub("skvideo.io")
    sys.modules["skvideo"].io = sys.modules["skvideo.io"]


def load_reference():
    """Returns (SelfCInvNet-instance-builder, GlobalVar, modules dict).  Raises if the mount is absent."""
    if not available():
        raise RuntimeError("/root/reference is not mounted (expected on the GPU box)")
    sys.dont_write_bytecode = True
    os.environ["PYTHONDONTWRITEBYTECODE"] = "1"
    _install_stubs()
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    saved = os.environ.get("CUDA_VISIBLE_DEVICES")
    import options.options as option           # noqa: E402  (reference module)
    import models.networks as networks         # noqa: E402
    from global_var import GlobalVar           # noqa: E402
    import models.modules.SelfC_GMM_arch_inv as arch  # noqa: E402

    def build_net(temporal_len: int):
        opt = option.parse(VID4_YAML, is_train=False)   # exports CUDA_VISIBLE_DEVICES=2 (F10)
        if saved is None:
            os.environ.pop("CUDA_VISIBLE_DEVICES", None)
        else:
            os.environ["CUDA_VISIBLE_DEVICES"] = saved
        opt = option.dict_to_nonedict(opt)
        GlobalVar.set_Temporal_LEN(temporal_len)
        net = networks.define_G(opt)
        net.eval()
        return net

    def inject_eps(net, eps):
        """Shim 2: same arithmetic as reparametrize (:412-417) with caller-provided eps."""
        def reparametrize(mu, logvar):
            import torch
            std = torch.exp(logvar)
            return eps.mul(std).add_(mu)
        net.stp_net.reparametrize = reparametrize

    return types.SimpleNamespace(build_net=build_net, inject_eps=inject_eps, GlobalVar=GlobalVar,
                                 arch=arch, option=option, networks=networks)
