"""Validation metrics on the GPU with the reference's definitions: `rgb_to_ycbcr` (data/util.py:239-245),
`calculate_psnr` (utils/util.py:198-221), `calculate_ssim` (utils/util.py:596-603 over `ssim` :361-488).  Same call
shapes and return types (lists of per-frame floats; `calculate_psnr` returns `inf` for the whole list when a frame
matches exactly, as the reference does), computed by csrc/metrics.cu instead of dozens of small torch launches."""
from __future__ import annotations

import math
from typing import Dict, List

import torch

from . import _lib
from .engine import _dev_check, _ptr, _stream

_WIN = {}


def _window(device) -> torch.Tensor:
    """_fspecial_gauss_1d(11, 1.5) (utils/util.py:361-375), in the same float32 arithmetic."""
    if device not in _WIN:
        coords = torch.arange(11).to(dtype=torch.float)
        coords -= 11 // 2
        g = torch.exp(-(coords ** 2) / (2 * 1.5 ** 2))
        g /= g.sum()
        _WIN[device] = g.to(device).contiguous()
    return _WIN[device]


def rgb_to_ycbcr(x: torch.Tensor) -> torch.Tensor:
    x = _dev_check(x)
    n, c, h, w = x.shape
    if c != 3:
        raise ValueError("rgb_to_ycbcr expects [N,3,H,W]")
    y = torch.empty((n, 1, h, w), dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        _lib.check(_lib.lib().selfc_rgb_to_y(_ptr(x), _ptr(y), n, h, w, _stream(x.device)), "rgb_to_y")
    return y


def _frame_sums(a: torch.Tensor, b: torch.Tensor, to_y: bool, want_ssim: bool):
    a, b = _dev_check(a), _dev_check(b)
    if a.shape != b.shape or a.dim() != 4:
        raise ValueError(f"metrics need two [N,C,H,W] tensors of the same shape, got {tuple(a.shape)} and {tuple(b.shape)}")
    n, c, h, w = a.shape
    sse = torch.zeros(n, dtype=torch.float64, device=a.device)
    ss = torch.zeros(n, dtype=torch.float64, device=a.device) if want_ssim else None
    with torch.cuda.device(a.device):
        _lib.check(_lib.lib().selfc_frame_metrics(_ptr(a), _ptr(b), n, c, h, w, 1 if to_y else 0, _ptr(_window(a.device)), _ptr(sse),
                                                  _ptr(ss), _stream(a.device)), "frame_metrics")
    ceff = 1 if to_y else c
    mse = (sse / float(ceff * h * w)).tolist()
    ssim = (ss / float(ceff * (h - 10) * (w - 10))).tolist() if want_ssim else None
    return mse, ssim


def _psnr_list(mse: List[float]):
    if any(m == 0 for m in mse):
        return float("inf")
    return [20.0 * math.log10(1.0 / math.sqrt(m)) for m in mse]


def calculate_psnr(img1s: torch.Tensor, img2s: torch.Tensor, to_y: bool = False):
    return _psnr_list(_frame_sums(img1s, img2s, to_y, False)[0])


def calculate_ssim(img1s: torch.Tensor, img2s: torch.Tensor, to_y: bool = False) -> List[float]:
    return _frame_sums(img1s, img2s, to_y, True)[1]


def clip_metrics(sr: torch.Tensor, gt: torch.Tensor, lr: torch.Tensor, lr_ref: torch.Tensor) -> Dict[str, list]:
    """The eight per-frame lists cal_metric accumulates (train.py:57-76): RGB and Y PSNR / SSIM of SR vs GT and LR vs LR_ref;
    the luma conversion is fused into the metric kernels."""
    out: Dict[str, list] = {}
    for tag, a, b in (("", sr, gt), ("lr_", lr, lr_ref)):
        for y in (False, True):
            mse, ssim = _frame_sums(a, b, y, True)
            p = _psnr_list(mse)
            sfx = "_y" if y else ""
            out[tag + "psnr" + sfx] = p if isinstance(p, list) else [p] * len(mse)
            out[tag + "ssim" + sfx] = ssim
    return out
