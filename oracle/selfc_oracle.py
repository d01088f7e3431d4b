"""CPU oracle for the SelfC-large 4x rescaling hot path.  TEST INFRASTRUCTURE ONLY.

A functional restatement (torch CPU fp32 ops over a plain ``state_dict``) of what
the reference network computes, written from the behaviour described in
SURVEY.md Appendix A.  Every function cites the reference lines it follows
(paths relative to /root/reference/codes).

Pinning status: the reference ships no tests or golden vectors for this path
(SURVEY.md F6), so the oracle is pinned against OUTPUTS OF THE REFERENCE ITSELF,
generated in the build container by ``oracle/make_golden.py`` (which imports the
unmodified reference modules from /root/reference) and committed under
``tests/golden/``.  ``tests/test_oracle_golden.py`` checks this file against
those vectors on every CPU test run.

Layouts here are the reference's: frames ``[B*T, C, H, W]`` fp32 NCHW, clips
``[B, C, T, h, w]``.
"""
from __future__ import annotations

import math
from collections import OrderedDict

import numpy as np
import torch
import torch.nn.functional as F

SCALE = 4
HF_DIM = 48          # 3 * 4 * 4
GMM_K = 5
GROWTH = 32          # D2DTInput gc
STP_C = 64


# --------------------------------------------------------------------------------------
# deterministic reference-layout weights (the bundled checkpoint is absent)
# --------------------------------------------------------------------------------------
def param_shapes() -> "OrderedDict[str, tuple]":
    """state_dict layout of SelfCInvNet for the vid4 YAML: 354 tensors.

    Order follows module registration order in the reference:
    models/modules/SelfC_GMM_arch_inv.py:433-448 (operations then stp_net),
    :300-344 (STPNet members), Subnet_constructor.py:98-106 (conv1..conv5),
    SelfC_GMM_arch_inv.py:258-263 (fc, proj1, proj2, proj3).
    """
    shapes: "OrderedDict[str, tuple]" = OrderedDict()

    def d2dt(prefix, cin, cout):
        for k in range(1, 5):
            shapes[f"{prefix}.conv{k}.weight"] = (GROWTH, cin + GROWTH * (k - 1), 1, 3, 3)
            shapes[f"{prefix}.conv{k}.bias"] = (GROWTH,)
        shapes[f"{prefix}.conv5.weight"] = (cout, cin + 4 * GROWTH, 3, 1, 1)
        shapes[f"{prefix}.conv5.bias"] = (cout,)

    def gagg(prefix, c=STP_C):
        shapes[f"{prefix}.fc.weight"] = (1, 1024)
        shapes[f"{prefix}.fc.bias"] = (1,)
        shapes[f"{prefix}.proj1.weight"] = (c, c, 1, 1)
        shapes[f"{prefix}.proj1.bias"] = (c,)
        shapes[f"{prefix}.proj2.weight"] = (c, c)
        shapes[f"{prefix}.proj2.bias"] = (c,)
        shapes[f"{prefix}.proj3.weight"] = (c, c)
        shapes[f"{prefix}.proj3.bias"] = (c,)

    for blk in range(1, 9):
        d2dt(f"operations.{blk}.F", HF_DIM, 3)
        d2dt(f"operations.{blk}.G", 3, HF_DIM)
        d2dt(f"operations.{blk}.H", 3, HF_DIM)
    d2dt("stp_net.local_m1", 3, STP_C)
    d2dt("stp_net.local_m2", STP_C, STP_C)
    gagg("stp_net.global_m1")
    gagg("stp_net.global_m2")
    for i in range(4):
        d2dt(f"stp_net.other_stp_modules.{2 * i}", STP_C, STP_C)
        gagg(f"stp_net.other_stp_modules.{2 * i + 1}")
    shapes["stp_net.tail_gmm.1.weight"] = (2 * STP_C, STP_C, 1, 1, 1)
    shapes["stp_net.tail_gmm.1.bias"] = (2 * STP_C,)
    shapes["stp_net.tail_gmm.3.weight"] = (4 * STP_C, 2 * STP_C, 1, 1, 1)
    shapes["stp_net.tail_gmm.3.bias"] = (4 * STP_C,)
    shapes["stp_net.tail_gmm.5.weight"] = (HF_DIM * GMM_K * 3, 4 * STP_C, 1, 1, 1)
    shapes["stp_net.tail_gmm.5.bias"] = (HF_DIM * GMM_K * 3,)
    return shapes


def make_state_dict(seed: int = 0, gain: float = 1.0) -> "OrderedDict[str, torch.Tensor]":
    """Seeded weights in the reference layout.

    U(-b, b) with b = gain / sqrt(fan_in), the PyTorch default Conv/Linear init the
    reference effectively keeps.  Generated with numpy's PCG64 so the same
    tensors are reproduced on any machine without the reference present.
    """
    rng = np.random.Generator(np.random.PCG64(seed))
    sd: "OrderedDict[str, torch.Tensor]" = OrderedDict()
    fan = {}
    for name, shape in param_shapes().items():
        base = name.rsplit(".", 1)[0]
        if name.endswith(".weight"):
            fan[base] = int(np.prod(shape[1:]))
        b = gain / math.sqrt(fan[base])
        sd[name] = torch.from_numpy(rng.uniform(-b, b, size=shape).astype(np.float32))
    return sd


def make_frames(b: int, t: int, hh: int, ww: int, seed: int = 1234) -> torch.Tensor:
    """Smooth synthetic 8-bit video ``[b*t, 3, hh, ww]`` in [0,1] (SURVEY 8d inputs)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    base = torch.from_numpy(rng.random((b * 3, 1, max(hh // 16, 2), max(ww // 16, 2)), dtype=np.float32))
    drift = torch.from_numpy(rng.random((b * 3, t, 1, 1), dtype=np.float32)) * 0.1
    up = F.interpolate(base, size=(hh, ww), mode="bicubic", align_corners=False)  # [b*3,1,hh,ww]
    clip = up + drift                                                            # [b*3,t,hh,ww]
    noise = torch.from_numpy(rng.standard_normal((b * 3, t, hh, ww), dtype=np.float32)) * 0.02
    clip = (clip + noise).clamp_(0, 1)
    clip = torch.round(clip * 255.0) / 255.0
    return clip.reshape(b, 3, t, hh, ww).transpose(1, 2).reshape(b * t, 3, hh, ww).contiguous()


def make_eps(b: int, t: int, h: int, w: int, seed: int) -> torch.Tensor:
    """Injected N(0,1) noise in the reference's layout ``[B,48,5,T,h,w]`` (:412-415), numpy PCG64."""
    rng = np.random.Generator(np.random.PCG64(seed))
    return torch.from_numpy(rng.standard_normal((b, HF_DIM, GMM_K, t, h, w), dtype=np.float32))


# --------------------------------------------------------------------------------------
# A.1 FrequencyAnalyzer  (models/modules/SelfC_GMM_arch_inv.py:46-82)
# --------------------------------------------------------------------------------------
def fa_forward(x: torch.Tensor, SCALE: int = SCALE) -> torch.Tensor:
    """``[N,3,H,W] -> [N,51,H/4,W/4]``: 4x4 box mean + unshuffled residual.

    HF channel = (sy*4+sx)*3 + c  (local PixelUnshuffle, :46-60, used at :75-78).
    SCALE=2 is the compression model's analyzer (SelfC_Codec_arch_inv.py:78-94): ``-> [N,15,H/2,W/2]``.
    """
    n, c, hh, ww = x.shape
    h, w = hh // SCALE, ww // SCALE
    blocks = x.reshape(n, c, h, SCALE, w, SCALE)
    # adaptive_avg_pool2d accumulates the window sequentially in row-major order, then divides
    # by the count; restated in that order so the result is bit-exact with nn.Upsample(mode="area").
    acc = torch.zeros(n, c, h, w, dtype=x.dtype)
    for sy in range(SCALE):
        for sx in range(SCALE):
            acc = acc + blocks[:, :, :, sy, :, sx]
    lf = acc / float(SCALE * SCALE)
    resid = blocks - lf[:, :, :, None, :, None]
    hf = resid.permute(0, 3, 5, 1, 2, 4).reshape(n, c * SCALE * SCALE, h, w)
    return torch.cat([lf, hf], dim=1)


def fa_reverse(z: torch.Tensor, SCALE: int = SCALE) -> torch.Tensor:
    """``[N,51,h,w] -> [N,3,4h,4w]``; HF channel = c*16 + sy*4 + sx (nn.PixelShuffle, :70,:79-82).

    NOT the inverse of fa_forward's channel order - replicated on purpose.
    SCALE=2: ``[N,15,h,w] -> [N,3,2h,2w]`` (SelfC_Codec_arch_inv.py:95-98).
    """
    n, _, h, w = z.shape
    lf = z[:, :3]
    hf = z[:, 3:].reshape(n, 3, SCALE, SCALE, h, w)
    out = lf[:, :, :, None, :, None] + hf.permute(0, 1, 4, 2, 5, 3)
    return out.reshape(n, 3, h * SCALE, w * SCALE)


# --------------------------------------------------------------------------------------
# F.3 HaarDownsampling  (models/modules/SelfC_arch_inv.py:44-84; Inv_arch.py:44-84 is the same class)
# --------------------------------------------------------------------------------------
_HAAR_SIGNS = ((1, 1, 1, 1), (1, -1, 1, -1), (1, 1, -1, -1), (1, -1, -1, 1))   # w_k over (a, b, c, d) = x[0,0], x[0,1], x[1,0], x[1,1]


def haar_forward(x: torch.Tensor) -> torch.Tensor:
    """``[N,C,H,W] -> [N,4C,H/2,W/2]``: grouped 2x2 stride-2 conv with the +-1 Haar weights (:50-60), ``/ 4.0``, then the
    ``reshape [N,C,4,..] -> transpose(1,2)`` of :70-73, i.e. output channel ``k*C + c``."""
    a, b, c, d = x[:, :, 0::2, 0::2], x[:, :, 0::2, 1::2], x[:, :, 1::2, 0::2], x[:, :, 1::2, 1::2]
    outs = [(((sa * a + sb * b) + sc * c) + sd * d) / 4.0 for sa, sb, sc, sd in _HAAR_SIGNS]
    return torch.cat(outs, dim=1)


def haar_reverse(z: torch.Tensor) -> torch.Tensor:
    """``[N,4C,h,w] -> [N,C,2h,2w]``: conv_transpose2d with the same weights, no scaling (:79-82)."""
    n, c4, h, w = z.shape
    cch = c4 // 4
    zk = [z[:, k * cch:(k + 1) * cch] for k in range(4)]
    y = torch.empty(n, cch, 2 * h, 2 * w, dtype=z.dtype)
    for pos, (dy, dx) in enumerate(((0, 0), (0, 1), (1, 0), (1, 1))):
        s = [_HAAR_SIGNS[k][pos] for k in range(4)]
        y[:, :, dy::2, dx::2] = ((s[0] * zk[0] + s[1] * zk[1]) + s[2] * zk[2]) + s[3] * zk[3]
    return y


# --------------------------------------------------------------------------------------
# A.3 D2DTInput dense block  (models/modules/Subnet_constructor.py:98-133)
# --------------------------------------------------------------------------------------
def _lrelu(x):
    return F.leaky_relu(x, 0.2)


def d2dt(sd, prefix: str, x: torch.Tensor, t: int) -> torch.Tensor:
    """``[B*T,Cin,h,w] -> [B*T,Cout,h,w]``; four (1,3,3) convs with dense concat then a
    (3,1,1) temporal conv zero-padded at clip ends (:126-130).  Concat order [X,x1..x4]."""
    bt, cin, h, w = x.shape
    b = bt // t
    v = x.reshape(b, t, cin, h, w).transpose(1, 2)
    feats = [v]
    for k in range(1, 5):
        y = F.conv3d(torch.cat(feats, 1), sd[f"{prefix}.conv{k}.weight"], sd[f"{prefix}.conv{k}.bias"],
                     padding=(0, 1, 1))
        feats.append(_lrelu(y))
    y = F.conv3d(torch.cat(feats, 1), sd[f"{prefix}.conv5.weight"], sd[f"{prefix}.conv5.bias"],
                 padding=(1, 0, 0))
    return y.transpose(1, 2).reshape(bt, -1, h, w)


# --------------------------------------------------------------------------------------
# A.2 InvBlockExp  (models/modules/SelfC_GMM_arch_inv.py:8-33), split 3/48, clamp=1
# --------------------------------------------------------------------------------------
def invblock_forward(sd, prefix: str, z: torch.Tensor, t: int) -> torch.Tensor:
    x1, x2 = z[:, :3], z[:, 3:]
    y1 = x1 + d2dt(sd, prefix + ".F", x2, t)                       # :25
    s = torch.sigmoid(d2dt(sd, prefix + ".H", y1, t)) * 2 - 1      # :26
    y2 = x2 * torch.exp(s) + d2dt(sd, prefix + ".G", y1, t)        # :27
    return torch.cat([y1, y2], 1)


def invblock_reverse(sd, prefix: str, z: torch.Tensor, t: int) -> torch.Tensor:
    x1, x2 = z[:, :3], z[:, 3:]
    s = torch.sigmoid(d2dt(sd, prefix + ".H", x1, t)) * 2 - 1      # :29
    y2 = (x2 - d2dt(sd, prefix + ".G", x1, t)) / torch.exp(s)      # :30
    y1 = x1 - d2dt(sd, prefix + ".F", y2, t)                       # :31
    return torch.cat([y1, y2], 1)


# --------------------------------------------------------------------------------------
# A.4 GlobalAgg  (models/modules/SelfC_GMM_arch_inv.py:257-285)
# --------------------------------------------------------------------------------------
def global_agg_weights(sd, prefix: str, x: torch.Tensor, t: int) -> torch.Tensor:
    """T x T mixing matrix ``[B,T,T]`` (rows softmaxed over the last index), :268-276."""
    bt, c, h, w = x.shape
    pooled = F.adaptive_avg_pool2d(x, (32, 32)).reshape(bt, c, 1024)
    d = (pooled @ sd[prefix + ".fc.weight"].t() + sd[prefix + ".fc.bias"]).reshape(bt // t, t, c)
    q = d @ sd[prefix + ".proj2.weight"].t() + sd[prefix + ".proj2.bias"]
    k = d @ sd[prefix + ".proj3.weight"].t() + sd[prefix + ".proj3.bias"]
    return torch.softmax(q @ k.transpose(1, 2) / c, dim=-1)


def global_agg(sd, prefix: str, x: torch.Tensor, t: int) -> torch.Tensor:
    """``out[b,t'] = x[b,t'] + sum_t proj1(x)[b,t] * W[b,t,t']``  (:266,:278-285)."""
    bt, c, h, w = x.shape
    wmat = global_agg_weights(sd, prefix, x, t)
    p = F.conv2d(x, sd[prefix + ".proj1.weight"], sd[prefix + ".proj1.bias"]).reshape(bt // t, t, c, h, w)
    mixed = torch.einsum("btchw,btu->buchw", p, wmat)
    return x + mixed.reshape(bt, c, h, w)


# --------------------------------------------------------------------------------------
# A.5 STPNet + A.6 sampler  (models/modules/SelfC_GMM_arch_inv.py:289-394,412-417)
# --------------------------------------------------------------------------------------
def stp_features(sd, lr: torch.Tensor, t: int) -> torch.Tensor:
    """``[B*T,3,h,w] -> [B*T,64,h,w]``  (:366-374)."""
    f = d2dt(sd, "stp_net.local_m1", lr, t)
    f = global_agg(sd, "stp_net.global_m1", f, t)
    f = d2dt(sd, "stp_net.local_m2", f, t)
    f = global_agg(sd, "stp_net.global_m2", f, t)
    for i in range(4):
        f = d2dt(sd, f"stp_net.other_stp_modules.{2 * i}", f, t)
        f = global_agg(sd, f"stp_net.other_stp_modules.{2 * i + 1}", f, t)
    return f


def gmm_head(sd, feat: torch.Tensor) -> torch.Tensor:
    """``[B*T,64,h,w] -> [B*T,720,h,w]``: lrelu -> 1x1 conv, three times (:336-344,:379)."""
    z = feat
    for idx in (1, 3, 5):
        wgt = sd[f"stp_net.tail_gmm.{idx}.weight"]
        z = F.conv2d(_lrelu(z), wgt.reshape(wgt.shape[0], wgt.shape[1], 1, 1), sd[f"stp_net.tail_gmm.{idx}.bias"])
    return z


def gmm_sample(params: torch.Tensor, eps: torch.Tensor, t: int) -> torch.Tensor:
    """Soft-GMM draw (:383-394).  ``params [B*T,720,h,w]`` with channel ``hf*15+k*3+j``
    (j=0 logit, 1 log-scale, 2 mean); softmax is over the 48 HF channels.
    ``eps`` in the reference's own layout ``[B,48,5,T,h,w]`` (:412-417).  Returns ``[B*T,48,h,w]``."""
    bt, _, h, w = params.shape
    b = bt // t
    p = params.reshape(b, t, HF_DIM, GMM_K, 3, h, w).permute(0, 2, 3, 4, 1, 5, 6)  # [B,48,5,3,T,h,w]
    pi = torch.softmax(p[:, :, :, 0], dim=1)
    std = torch.exp(torch.clamp(p[:, :, :, 1], -7, 7))
    mu = p[:, :, :, 2]
    v = (pi * (eps * std + mu)).sum(2)                                              # [B,48,T,h,w]
    return v.transpose(1, 2).reshape(bt, HF_DIM, h, w)


# --------------------------------------------------------------------------------------
# A.7 quantisation  (models/modules/Quantization.py:4-17)
# --------------------------------------------------------------------------------------
def quantize(x: torch.Tensor) -> torch.Tensor:
    return torch.round(torch.clamp(x, 0, 1) * 255.0) / 255.0


def quantize_u8(x: torch.Tensor) -> torch.Tensor:
    return torch.round(torch.clamp(x, 0, 1) * 255.0).to(torch.uint8)


# --------------------------------------------------------------------------------------
# SelfCInvNet.forward  (models/modules/SelfC_GMM_arch_inv.py:450-490)
# --------------------------------------------------------------------------------------
def net_down(sd, x: torch.Tensor, t: int, return_stages: bool = False):
    """rev=False: ``[B*T,3,H,W] -> [B*T,51,h,w]`` (FA then blocks 1..8)."""
    z = fa_forward(x)
    stages = [z]
    for blk in range(1, 9):
        z = invblock_forward(sd, f"operations.{blk}", z, t)
        stages.append(z)
    return (z, stages) if return_stages else z


def net_up(sd, lr: torch.Tensor, eps: torch.Tensor, t: int, return_stages: bool = False):
    """rev=True: quantised LR ``[B*T,3,h,w]`` + injected eps -> ``(HR [B*T,3,H,W], hf [B*T,48,h,w])``."""
    feat = stp_features(sd, lr[:, :3], t)
    params = gmm_head(sd, feat)
    hf = gmm_sample(params, eps, t)
    z = torch.cat([lr[:, :3], hf], 1)
    stages = [z]
    for blk in range(8, 0, -1):
        z = invblock_reverse(sd, f"operations.{blk}", z, t)
        stages.append(z)
    hr = fa_reverse(z)
    if return_stages:
        return hr, hf, {"feat": feat, "params": params, "stages": stages}
    return hr, hf


def frames_from_u8(img: np.ndarray) -> torch.Tensor:
    """Decoded 8-bit frames ``[n,H,W,3]`` (cv2 order B,G,R) -> ``[n,3,H,W]`` RGB fp32 in [0,1].

    read_img1's ``astype(np.float32) / 255.`` (data/util.py:103-115) followed by the dataset's ``[:, :, [2, 1, 0]]`` and
    ``np.transpose(..., (2, 0, 1))`` (data/LQGTVID_dataset.py:150-154), per frame."""
    x = img.astype(np.float32) / 255.
    x = x[..., [2, 1, 0]]
    return torch.from_numpy(np.ascontiguousarray(np.transpose(x, (0, 3, 1, 2)))).float()


def frames_to_u8(x: torch.Tensor) -> np.ndarray:
    """``[n,3,H,W]`` RGB fp32 -> 8-bit frames ``[n,H,W,3]`` (B,G,R): tensor2img per frame (utils/util.py:104-133):
    clamp to [0,1], (t-0)/(1-0), channel flip, CHW->HWC, ``(img * 255.0).round()`` (numpy: half to even), astype(uint8)."""
    t = x.float().clamp(0, 1)
    t = (t - 0) / (1 - 0)
    a = np.transpose(t.numpy()[:, [2, 1, 0]], (0, 2, 3, 1))
    return (a * 255.0).round().astype(np.uint8)


def rescale_u8(sd, img: np.ndarray, eps: torch.Tensor, t: int):
    """rescale() on 8-bit frames: (LR frames, reconstructed HR frames), both uint8 [n,·,·,3] BGR."""
    lr, hr = rescale(sd, frames_from_u8(img), eps, t)
    return frames_to_u8(lr), frames_to_u8(hr)


def rescale(sd, x: torch.Tensor, eps: torch.Tensor, t: int):
    """down -> 8-bit quantise -> up: one 'step' of the benchmark (models/SelfC_model.py:213-233)."""
    z = net_down(sd, x, t)
    lr = quantize(z[:, :3])
    hr, _ = net_up(sd, lr, eps, t)
    return lr, hr


# --------------------------------------------------------------------------------------
# A.9 metrics  (data/util.py:239-245, utils/util.py:198-221, :361-439, :597-605)
# --------------------------------------------------------------------------------------
def rgb_to_y(x: torch.Tensor) -> torch.Tensor:
    y = x[:, 0:1] * 65.481 + x[:, 1:2] * 128.553 + x[:, 2:3] * 24.966 + 16.0
    return y / 255.0


def psnr_frames(a: torch.Tensor, b: torch.Tensor):
    """Per-frame PSNR on ``[N,1,H,W]``; returns inf for the whole list if any mse is 0 (:214-219)."""
    out = []
    for i in range(a.shape[0]):
        mse = torch.mean((a[i] - b[i]) ** 2.0)
        if mse == 0:
            return float("inf")
        out.append(20.0 * torch.log10(1.0 / torch.sqrt(mse)).item())
    return out


def _gauss_win(size=11, sigma=1.5):
    c = torch.arange(size, dtype=torch.float) - size // 2
    g = torch.exp(-(c ** 2) / (2 * sigma ** 2))
    return (g / g.sum()).reshape(1, 1, 1, size)


def ssim_frames(a: torch.Tensor, b: torch.Tensor):
    """Per-frame SSIM on ``[N,1,H,W]``, 11-tap sigma=1.5 separable Gaussian, valid conv, data_range=1."""
    win = _gauss_win()
    c1, c2 = 0.01 ** 2, 0.03 ** 2

    def blur(v):
        v = F.conv2d(v, win)
        return F.conv2d(v, win.transpose(2, 3))

    out = []
    for i in range(a.shape[0]):
        x, y = a[i:i + 1], b[i:i + 1]
        mu1, mu2 = blur(x), blur(y)
        s1 = blur(x * x) - mu1 * mu1
        s2 = blur(y * y) - mu2 * mu2
        s12 = blur(x * y) - mu1 * mu2
        cs = (2 * s12 + c2) / (s1 + s2 + c2)
        m = ((2 * mu1 * mu2 + c1) / (mu1 * mu1 + mu2 * mu2 + c1)) * cs
        out.append(m.mean().item())
    return out


# --------------------------------------------------------------------------------------
# LR_ref: DUF Gaussian blur-downsample  (models/Guassian.py:7-52), scale 4
# --------------------------------------------------------------------------------------
def gaussian_kernel_13(sigma: float = 1.6) -> torch.Tensor:
    """13x13 kernel = scipy.ndimage.gaussian_filter of a dirac (truncate 4 sigma, reflect)."""
    import scipy.ndimage

    d = np.zeros((13, 13))
    d[6, 6] = 1
    return torch.from_numpy(scipy.ndimage.gaussian_filter(d, sigma)).float()


def gaussian_downsample(x: torch.Tensor) -> torch.Tensor:
    """``[N,3,H,W] -> [N,3,H/4,W/4]``: reflect-pad 14, 13x13 sigma 1.6, stride 4, crop 2."""
    n, c, hh, ww = x.shape
    v = F.pad(x.reshape(-1, 1, hh, ww), [14, 14, 14, 14], mode="reflect")
    v = F.conv2d(v, gaussian_kernel_13()[None, None], stride=4)[:, :, 2:-2, 2:-2]
    return v.reshape(n, c, v.shape[2], v.shape[3])


# --------------------------------------------------------------------------------------
# a13 training step (models/SelfC_model.py:148-183, models/modules/loss.py:5-21, Quantization.py:4-17) -- the autograd
# reference for the backward kernels of the next round; pinned by tests/golden/train_t3.npz
# --------------------------------------------------------------------------------------
def quantize_ste(x: torch.Tensor) -> torch.Tensor:
    """Quant.forward / backward (Quantization.py:4-17): clamp + round to the 1/255 grid, gradient = identity everywhere
    (the clamp sits inside the autograd.Function, so the straight-through estimator also passes outside [0,1])."""
    return x + (quantize(x) - x).detach()


def train_losses(sd, x: torch.Tensor, ref_l: torch.Tensor, eps: torch.Tensor, t: int):
    """optimize_parameters' forward (SelfC_model.py:152-170) with the training YAML's settings (l2 forward fit, l1
    = Charbonnier eps 1e-6 backward reconstruction, lambda 1/1, loss_c = 0): returns (loss, l_forw_fit, l_back_rec)."""
    out = net_down(sd, x, t)
    lr_pre = out[:, :3]
    l_forw = ((lr_pre - ref_l.detach()) ** 2).mean()
    hr, _ = net_up(sd, quantize_ste(lr_pre), eps, t)
    d = x - hr[:, :3]
    l_back = torch.sqrt(d * d + 1e-6).mean()
    loss = (l_forw + l_back + 0.0) * 144 * 144 * 3
    return loss, l_forw, l_back


def train_grads(sd, x: torch.Tensor, ref_l: torch.Tensor, eps: torch.Tensor, t: int):
    """loss.backward() of the step above: {name: grad} for every parameter, plus the three loss values."""
    leaf = {k: v.detach().clone().requires_grad_(True) for k, v in sd.items()}
    loss, l_forw, l_back = train_losses(leaf, x, ref_l, eps, t)
    loss.backward()
    return {k: v.grad for k, v in leaf.items()}, float(loss.detach()), float(l_forw.detach()), float(l_back.detach())


# --------------------------------------------------------------------------------------
# counter-based noise of the CUDA path (selfc_b200/csrc/common.cuh: philox_normal) - numpy restatement
# --------------------------------------------------------------------------------------
def philox4x32_10(c, k):
    """Philox4x32-10 (Salmon et al., SC'11): c uint32[...,4] counters, k uint32[...,2] keys -> uint32[...,4]."""
    c = [c[..., i].astype(np.uint64) for i in range(4)]
    k0 = k[..., 0].astype(np.uint64)
    k1 = k[..., 1].astype(np.uint64)
    M0, M1, W0, W1, MASK = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57), np.uint64(0x9E3779B9), np.uint64(0xBB67AE85), np.uint64(0xFFFFFFFF)
    for _ in range(10):
        p0 = M0 * c[0]
        p1 = M1 * c[2]
        n0 = ((p1 >> np.uint64(32)) ^ c[1] ^ k0) & MASK
        n1 = p1 & MASK
        n2 = ((p0 >> np.uint64(32)) ^ c[3] ^ k1) & MASK
        n3 = p0 & MASK
        c = [n0, n1, n2, n3]
        k0 = (k0 + W0) & MASK
        k1 = (k1 + W1) & MASK
    return np.stack(c, -1).astype(np.uint32)


def philox_normal4(g: np.ndarray, seed: int, offset: int) -> np.ndarray:
    """The four normals of Philox call `g` of stream (seed, offset) (common.cuh: philox_normal4): counter
    (g_lo, g_hi, off_lo, off_hi), key (seed_lo, seed_hi); Box-Muller on words (0,1) -> [cos, sin] and (2,3) -> [cos, sin],
    in float32 like the device code.  Returns float32[..., 4]."""
    g = np.asarray(g, dtype=np.uint64)
    ctr = np.stack([g & np.uint64(0xFFFFFFFF), g >> np.uint64(32),
                    np.full_like(g, offset & 0xFFFFFFFF), np.full_like(g, (offset >> 32) & 0xFFFFFFFF)], -1).astype(np.uint32)
    key = np.stack([np.full_like(g, seed & 0xFFFFFFFF), np.full_like(g, (seed >> 32) & 0xFFFFFFFF)], -1).astype(np.uint32)
    r = philox4x32_10(ctr, key)
    scale = np.float32(2.3283064365386963e-10)
    out = []
    for h in range(2):
        u1 = (r[..., 2 * h].astype(np.float32) + np.float32(1.0)) * scale
        u2 = r[..., 2 * h + 1].astype(np.float32) * scale
        rad = np.sqrt(np.float32(-2.0) * np.log(u1).astype(np.float32)).astype(np.float32)
        ang = np.float64(np.pi) * (np.float32(2.0) * u2).astype(np.float64)
        out.append((rad * np.cos(ang).astype(np.float32)).astype(np.float32))
        out.append((rad * np.sin(ang).astype(np.float32)).astype(np.float32))
    return np.stack(out, -1)


def philox_eps(b: int, t: int, h: int, w: int, seed: int, offset: int) -> np.ndarray:
    """The device noise tensor eps [b,48,5,t,h,w]: Philox call g = linear index of (b, hf//4, k, t, pixel) in [b,12,5,t,h*w]
    gives hf = 4*(hf//4) + 0..3 (DESIGN.md "Noise")."""
    g = np.arange(b * 12 * 5 * t * h * w, dtype=np.uint64)
    n4 = philox_normal4(g, seed, offset).reshape(b, 12, 5, t, h, w, 4)
    return np.ascontiguousarray(np.moveaxis(n4, -1, 2)).reshape(b, 48, 5, t, h, w)
