"""Drop-in for the reference's `models.modules.SelfC_GMM_arch_inv.SelfCInvNet` (the SelfC-large rescaler).

Same constructor signature, same `forward(x, rev, cal_jacobian, lr_before_distor)` returns, same module tree and
therefore the same `state_dict` keys / shapes -- it strict-loads a reference checkpoint -- but
`forward` runs the hand-written sm_100a kernels of libselfc_b200 through the C-ABI instead of PyTorch ops.

Reference lines mirrored (paths relative to /root/reference/codes):
  SelfCInvNet            models/modules/SelfC_GMM_arch_inv.py:432-494
  InvBlockExp            :8-41          FrequencyAnalyzer  :62-82
  GlobalAgg              :257-285       STPNet             :289-430
  D2DTInput              models/modules/Subnet_constructor.py:98-133

The sub-modules below only HOLD parameters in the reference's names and shapes (module construction order
matches the reference, so `torch.manual_seed(s); define_G(opt)` yields the same initial weights);
the arithmetic lives in csrc/.  There is no PyTorch fallback: CPU tensors raise.
"""
from __future__ import annotations

import os
from typing import Optional

import torch
import torch.nn as nn

from . import engine as _engine
from .global_var import GlobalVar


class D2DTInput(nn.Module):
    """Parameter holder for the dense block (Subnet_constructor.py:98-106)."""

    def __init__(self, channel_in, channel_out, init="xavier", gc=32, bias=True, INN_init=True, is_res=False):
        super().__init__()
        self.conv1 = nn.Conv3d(channel_in, gc, (1, 3, 3), 1, (0, 1, 1), bias=bias)
        self.conv2 = nn.Conv3d(channel_in + gc, gc, (1, 3, 3), 1, (0, 1, 1), bias=bias)
        self.conv3 = nn.Conv3d(channel_in + 2 * gc, gc, (1, 3, 3), 1, (0, 1, 1), bias=bias)
        self.conv4 = nn.Conv3d(channel_in + 3 * gc, gc, (1, 3, 3), 1, (0, 1, 1), bias=bias)
        self.conv5 = nn.Conv3d(channel_in + 4 * gc, channel_out, (3, 1, 1), 1, (1, 0, 0), bias=bias)
        # The reference's xavier*0.1 / zero-init helpers only touch Conv2d/Linear/BatchNorm2d
        # (module_util.py:7-44) and are a no-op on these Conv3d layers: default init is kept.


def subnet(net_structure, init="xavier"):
    """Subnet_constructor.py:719-788, D2DTNet branch only (the one SelfC-large selects)."""
    def constructor(channel_in, channel_out):
        if net_structure == "D2DTNet":
            return D2DTInput(channel_in, channel_out, init)
        raise NotImplementedError(f"selfc_b200 implements subnet_type 'D2DTNet' only (got {net_structure!r})")
    return constructor


class InvBlockExp(nn.Module):
    def __init__(self, subnet_constructor, channel_num, channel_split_num, clamp=1.0):
        super().__init__()
        self.split_len1 = channel_split_num
        self.split_len2 = channel_num - channel_split_num
        self.clamp = clamp
        self.F = subnet_constructor(self.split_len2, self.split_len1)
        self.G = subnet_constructor(self.split_len1, self.split_len2)
        self.H = subnet_constructor(self.split_len1, self.split_len2)


class FrequencyAnalyzer(nn.Module):
    """k x k box-mean LF + pixel-unshuffled residual HF; usable on its own.  k=4: the rescaler's
    (SelfC_GMM_arch_inv.py:62-82); k=2: the compression model's (SelfC_Codec_arch_inv.py:78-98, its default)."""

    def __init__(self, channel_in=3, k=4):
        super().__init__()
        if channel_in != 3 or k not in (2, 4):
            raise NotImplementedError("FrequencyAnalyzer is built for 3 channels and k in {2, 4}")
        self.k = k

    def forward(self, x, rev=False):
        return _engine.fa_reverse(x, self.k) if rev else _engine.fa_forward(x, self.k)


class HaarDownsampling(nn.Module):
    """Orthogonal 2x2 Haar split of `model: SelfC` / IRN (SelfC_arch_inv.py:44-84, Inv_arch.py:44-84).  `haar_weights` is kept
    as the reference's frozen parameter so that state_dicts stay interchangeable; the kernels hard-code its +-1 pattern."""

    def __init__(self, channel_in):
        super().__init__()
        self.channel_in = channel_in
        w = torch.ones(4, 1, 2, 2)
        w[1, 0, 0, 1] = -1
        w[1, 0, 1, 1] = -1
        w[2, 0, 1, 0] = -1
        w[2, 0, 1, 1] = -1
        w[3, 0, 1, 0] = -1
        w[3, 0, 0, 1] = -1
        self.haar_weights = nn.Parameter(torch.cat([w] * channel_in, 0), requires_grad=False)
        self.last_jac = 0.0

    def forward(self, x, rev=False):
        import math
        elements = x.shape[1] * x.shape[2] * x.shape[3]
        if not rev:
            self.last_jac = elements / 4 * math.log(1 / 16.)
            return _engine.haar_forward(x)
        self.last_jac = elements / 4 * math.log(16.)
        return _engine.haar_reverse(x)

    def jacobian(self, x, rev=False):
        return self.last_jac


class GlobalAgg(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.fc = nn.Linear(32 * 32, 1)
        self.proj1 = nn.Conv2d(c, c, 1, 1, 0)
        self.proj2 = nn.Linear(c, c)
        self.proj3 = nn.Linear(c, c)


class STPNet(nn.Module):
    def __init__(self, opt):
        super().__init__()
        self.global_module = opt["global_module"]
        self.stp_blk_num = opt["stp_blk_num"]
        self.fh_loss = opt["fh_loss"]
        self.scale = opt["scale"]
        self.K = opt["gmm_k"]
        if self.global_module != "nonlocal" or self.fh_loss != "gmm" or self.K != 5 or self.scale != 4 or self.stp_blk_num != 6:
            raise NotImplementedError(
                "selfc_b200 implements the SelfC-large prior only: global_module=nonlocal, fh_loss=gmm, gmm_k=5, "
                f"scale=4, stp_blk_num=6 (got {self.global_module}, {self.fh_loss}, {self.K}, {self.scale}, {self.stp_blk_num})")
        self.stp_blk_num = self.stp_blk_num - 2
        c = 64
        self.local_m1 = D2DTInput(3, c, INN_init=False)
        self.local_m2 = D2DTInput(c, c, INN_init=False)
        self.global_m1 = GlobalAgg(c)
        self.global_m2 = GlobalAgg(c)
        mods = []
        for _ in range(self.stp_blk_num):
            mods += [D2DTInput(c, c, INN_init=False), GlobalAgg(c)]
        self.other_stp_modules = nn.Sequential(*mods)
        self.hf_dim = 3 * (self.scale ** 2)
        self.tail_gmm = nn.Sequential(
            nn.LeakyReLU(negative_slope=0.2, inplace=True),
            nn.Conv3d(c, c * 2, 1, 1, 0, bias=True),
            nn.LeakyReLU(negative_slope=0.2, inplace=True),
            nn.Conv3d(c * 2, c * 4, 1, 1, 0, bias=True),
            nn.LeakyReLU(negative_slope=0.2, inplace=True),
            nn.Conv3d(c * 4, self.hf_dim * self.K * 3, 1, 1, 0, bias=True))
        self.gmm_v = None

    def sample(self):
        return self.gmm_v


class SelfCInvNet(nn.Module):
    def __init__(self, opt, channel_in=3, channel_out=3, subnet_type="D2DTNet", block_num=(4, 4), down_num=2):
        super().__init__()
        if channel_in != 3 or channel_out != 3:
            raise NotImplementedError("selfc_b200 implements the 3-channel RGB rescaler only")
        n_blocks = sum(block_num[i] for i in range(down_num))
        if n_blocks != 8:
            raise NotImplementedError(f"selfc_b200 implements the 8-block SelfC-large rescaler (got {n_blocks} blocks)")
        subnet_constructor = subnet(subnet_type, "xavier")
        operations = [FrequencyAnalyzer(channel_in)]
        current_channel = channel_in * 17
        for i in range(down_num):
            for _ in range(block_num[i]):
                operations.append(InvBlockExp(subnet_constructor, current_channel, channel_out))
        self.operations = nn.ModuleList(operations)
        self.stp_net = STPNet(opt)
        # new, out-of-band knobs (absent = reference behaviour: fp32)
        self.precision = (opt.get("precision") if hasattr(opt, "get") else None) or os.environ.get("SELFC_B200_PRECISION", "fp32")
        self.noise_seed = 0
        self.noise_offset = 0
        self._eps_override: Optional[torch.Tensor] = None
        self._engines = {}
        self._named = None

    # ---- noise control (the reference draws unseeded CUDA noise, :412-417) ---------------------------------
    def set_noise(self, seed: int, offset: int = 0):
        self.noise_seed, self.noise_offset = int(seed), int(offset)

    def inject_eps(self, eps: Optional[torch.Tensor]):
        """Use this [B,48,5,T,h,w] tensor instead of Philox noise for subsequent rev=True calls (None = off)."""
        self._eps_override = eps

    def set_precision(self, mode: str):
        _engine.parse_mode(mode)
        self.precision = mode
        self._engines.clear()

    # ---- engine plumbing ----------------------------------------------------------------------------------
    def _engine_for(self, device: torch.device) -> "_engine.Engine":
        if device.type != "cuda":
            raise RuntimeError("selfc_b200.SelfCInvNet runs on CUDA (sm_100a) only; there is no CPU / PyTorch fallback")
        key = (device.index if device.index is not None else torch.cuda.current_device(), self.precision)
        eng = self._engines.get(key)
        if eng is None:
            eng = _engine.Engine(torch.device("cuda", key[0]), self.precision)
            self._engines[key] = eng
        if self._named is None:
            own = dict(self.named_parameters())
            self._named = [(n, own[n]) for n in _engine.PARAM_NAMES]
        eng.sync_params(self._named)
        return eng

    @property
    def engine(self) -> "_engine.Engine":
        """The C-ABI engine holding this module's current weights on its device (8-bit frame calls)."""
        return self._engine_for(next(self.parameters()).device)

    def rescale_u8(self, frames_bgr: torch.Tensor):
        """uint8 [B*T,H,W,3] (cv2 layout) -> (LR frames, reconstructed HR frames), uint8, the same layout: what the test loop's
        read_img1 -> forward -> Quantization -> forward(rev=True) -> tensor2img chain produces (models/SelfC_model.py:213-233)."""
        t = GlobalVar.get_Temporal_LEN()
        if t is None:
            raise RuntimeError("GlobalVar.set_Temporal_LEN(T) must be called before rescale_u8")
        out = self.engine.rescale_u8(frames_bgr, t, seed=self.noise_seed, offset=self.noise_offset, eps=self._eps_override)
        self.noise_offset += 1
        return out

    def _apply(self, fn, *a, **k):   # .to()/.cuda()/.half(): parameter storage changes -> rebuild the name list
        self._named = None
        return super()._apply(fn, *a, **k)

    def load_state_dict(self, *a, **k):
        self._named = None
        return super().load_state_dict(*a, **k)

    def forward(self, x, rev=False, cal_jacobian=False, lr_before_distor=None):
        if cal_jacobian:
            raise NotImplementedError("cal_jacobian is not used by the rescaling path and is not implemented")
        t = GlobalVar.get_Temporal_LEN()
        if t is None:
            raise RuntimeError("GlobalVar.set_Temporal_LEN(T) must be called before forward (the reference's dataset does it)")
        eng = self._engine_for(x.device)            # raises for CPU tensors: there is no fallback
        if self.training and torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()):
            raise RuntimeError("selfc_b200.SelfCInvNet.forward builds no autograd graph: loss.backward() through it cannot work. "
                               "Train through SelfCModel.optimize_parameters / selfc_b200.train.Trainer (the CUDA training "
                               "step), or call forward under torch.no_grad() / in eval() mode.")
        if not rev:
            out, _, _ = eng.down(x, t, want_out51=True, want_u8=False, want_q=False)
            return out, out.new_zeros(())          # loss_c = out.mean() * 0  (:468)
        lr = x[:, 0:3]
        hr, hf = eng.up(lr, t, eps=self._eps_override, seed=self.noise_seed, offset=self.noise_offset, want_hf=True)
        self.noise_offset += 1
        bt, c, h, w = hf.shape
        self.stp_net.gmm_v = hf.reshape(bt // t, t, c, h, w).transpose(1, 2)
        return hr, hf
