"""Host-side driver of libselfc_b200: owns a selfc_ctx (packed-weight cache) and the activation workspace for one
device, and exposes the path as tensor-in / tensor-out calls.  PyTorch is used for device memory and streams only.

There is no CPU path: every call requires CUDA tensors and the built sm_100a library.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Optional, Sequence, Tuple

import torch

from . import _lib
from ._lib import MODE_BF16, MODE_BF16X3, MODE_FP32, NUM_PARAMS

HF_DIM = 48
GMM_K = 5


def param_names() -> Sequence[str]:
    """state_dict keys of SelfCInvNet in the canonical order the C-ABI expects."""
    names = []

    def d2dt(prefix):
        for k in range(1, 6):
            names.append(f"{prefix}.conv{k}.weight")
            names.append(f"{prefix}.conv{k}.bias")

    def gagg(prefix):
        for m in ("fc", "proj1", "proj2", "proj3"):
            names.append(f"{prefix}.{m}.weight")
            names.append(f"{prefix}.{m}.bias")

    for blk in range(1, 9):
        for sub in "FGH":
            d2dt(f"operations.{blk}.{sub}")
    d2dt("stp_net.local_m1")
    d2dt("stp_net.local_m2")
    gagg("stp_net.global_m1")
    gagg("stp_net.global_m2")
    for i in range(4):
        d2dt(f"stp_net.other_stp_modules.{2 * i}")
        gagg(f"stp_net.other_stp_modules.{2 * i + 1}")
    for idx in (1, 3, 5):
        names.append(f"stp_net.tail_gmm.{idx}.weight")
        names.append(f"stp_net.tail_gmm.{idx}.bias")
    assert len(names) == NUM_PARAMS
    return names


PARAM_NAMES = tuple(param_names())
PARAM_INDEX = {n: i for i, n in enumerate(PARAM_NAMES)}


def _ptr(t: Optional[torch.Tensor]) -> C.c_void_p:
    return C.c_void_p(0 if t is None else t.data_ptr())


def _stream(device: torch.device) -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def parse_mode(mode) -> int:
    if mode in (MODE_FP32, "fp32", "float32", None):
        return MODE_FP32
    if mode in (MODE_BF16, "bf16", "bfloat16"):
        return MODE_BF16
    if mode in (MODE_BF16X3, "bf16x3", "fp32_tc"):
        return MODE_BF16X3
    raise ValueError(f"unknown selfc_b200 precision mode {mode!r} (use 'fp32', 'bf16x3' or 'bf16')")


class Engine:
    """One selfc_ctx + workspace on one CUDA device."""

    def __init__(self, device: torch.device, mode="fp32"):
        device = torch.device(device)
        if device.type != "cuda":
            raise RuntimeError("selfc_b200 runs on CUDA (sm_100a) only; there is no CPU fallback")
        self.device = torch.device("cuda", device.index if device.index is not None else torch.cuda.current_device())
        self.mode = parse_mode(mode)
        self._L = _lib.lib()
        h = C.c_void_p()
        _lib.check(self._L.selfc_ctx_create(C.byref(h), self.device.index, self.mode), "ctx_create")
        self._ctx = h
        self._ws: Optional[torch.Tensor] = None
        self._weights_key = None
        self._keep = None   # keeps fp32 contiguous copies of the parameters alive while packing

    def close(self):
        if getattr(self, "_ctx", None) is not None and self._ctx.value:
            self._L.selfc_ctx_destroy(self._ctx)
            self._ctx = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- weights ----------------------------------------------------------------------------------------
    def load_state(self, state: Dict[str, torch.Tensor]) -> None:
        """(Re)pack weights from a reference-layout state_dict (keys may carry a 'module.' prefix)."""
        tensors = []
        for name in PARAM_NAMES:
            t = state.get(name)
            if t is None:
                t = state.get("module." + name)
            if t is None:
                raise KeyError(f"state_dict is missing {name}")
            t = t.detach().to(device=self.device, dtype=torch.float32).contiguous()
            tensors.append(t)
        arr = (C.c_void_p * NUM_PARAMS)(*[t.data_ptr() for t in tensors])
        with torch.cuda.device(self.device):
            _lib.check(self._L.selfc_ctx_load_weights(self._ctx, arr, NUM_PARAMS, _stream(self.device)), "load_weights")
        self._keep = tensors
        self._shapes = {name: tuple(t.shape) for name, t in zip(PARAM_NAMES, tensors)}

    def invalidate(self) -> None:
        """Forget the packed-weight cache key: the next sync_params re-packs (used after an in-place optimiser step, which
        writes the parameters through raw pointers and therefore does not bump their version counters)."""
        self._weights_key = None

    def sync_params(self, named_params: Sequence[Tuple[str, torch.Tensor]]) -> None:
        """Re-pack iff any parameter's storage or version changed since the last call."""
        key = tuple((p.data_ptr(), p._version) for _, p in named_params)
        if key != self._weights_key:
            self.load_state({n: p for n, p in named_params})
            self._weights_key = key

    # ---- workspace ---------------------------------------------------------------------------------------
    def _workspace(self, B: int, T: int, h: int, w: int) -> torch.Tensor:
        need = int(self._L.selfc_workspace_bytes(self._ctx, B, T, h, w))
        if need == 0:
            raise RuntimeError(f"selfc_b200: bad clip shape B={B} T={T} h={h} w={w}")
        if self._ws is None or self._ws.numel() < need:
            self._ws = None
            self._ws = torch.empty(need, dtype=torch.uint8, device=self.device)
        return self._ws

    @staticmethod
    def _clip_dims(x: torch.Tensor, T: int) -> Tuple[int, int, int]:
        bt, _, hh, ww = x.shape
        if T is None or T < 1 or bt % T != 0:
            raise ValueError(f"batch of {bt} frames is not a multiple of the clip length T={T} (GlobalVar.set_Temporal_LEN)")
        return bt // T, hh, ww

    def _check_in(self, x: torch.Tensor, what: str) -> torch.Tensor:
        if not x.is_cuda or x.device != self.device:
            raise RuntimeError(f"selfc_b200: {what} must live on {self.device} (got {x.device}); there is no CPU path")
        return x.to(torch.float32).contiguous()

    # ---- the path ----------------------------------------------------------------------------------------
    def down(self, hr: torch.Tensor, T: int, want_out51: bool = True, want_u8: bool = True, want_q: bool = True):
        """SelfCInvNet.forward(rev=False) + Quantization.  hr [B*T,3,H,W] -> (out51|None, lr_u8|None, lr_q|None)."""
        hr = self._check_in(hr, "hr")
        B, H, W = self._clip_dims(hr, T)
        if hr.shape[1] != 3 or H % 4 or W % 4:
            raise ValueError(f"expected [B*T,3,H,W] with H,W multiples of 4, got {tuple(hr.shape)}")
        h, w = H // 4, W // 4
        ws = self._workspace(B, T, h, w)
        out51 = torch.empty((B * T, 51, h, w), dtype=torch.float32, device=self.device) if want_out51 else None
        lr_u8 = torch.empty((B * T, 3, h, w), dtype=torch.uint8, device=self.device) if want_u8 else None
        lr_q = torch.empty((B * T, 3, h, w), dtype=torch.float32, device=self.device) if want_q else None
        with torch.cuda.device(self.device):
            _lib.check(self._L.selfc_down(self._ctx, _ptr(hr), _ptr(out51), _ptr(lr_u8), _ptr(lr_q), B, T, H, W,
                                          _ptr(ws), ws.numel(), _stream(self.device)), "down")
        return out51, lr_u8, lr_q

    def up(self, lr: torch.Tensor, T: int, eps: Optional[torch.Tensor] = None, seed: int = 0, offset: int = 0,
           want_hf: bool = True):
        """SelfCInvNet.forward(rev=True).  lr [B*T,3,h,w] -> (hr [B*T,3,4h,4w], recon_hf [B*T,48,h,w]|None)."""
        lr = self._check_in(lr, "lr")
        B, h, w = self._clip_dims(lr, T)
        if lr.shape[1] != 3:
            raise ValueError(f"expected [B*T,3,h,w], got {tuple(lr.shape)}")
        if eps is not None:
            eps = self._check_in(eps, "eps")
            if tuple(eps.shape) != (B, HF_DIM, GMM_K, T, h, w):
                raise ValueError(f"eps must be [B,48,5,T,h,w]={(B, HF_DIM, GMM_K, T, h, w)}, got {tuple(eps.shape)}")
        ws = self._workspace(B, T, h, w)
        hr = torch.empty((B * T, 3, 4 * h, 4 * w), dtype=torch.float32, device=self.device)
        hf = torch.empty((B * T, HF_DIM, h, w), dtype=torch.float32, device=self.device) if want_hf else None
        with torch.cuda.device(self.device):
            _lib.check(self._L.selfc_up(self._ctx, _ptr(lr), _ptr(eps), seed & (2 ** 64 - 1), offset & (2 ** 64 - 1),
                                        _ptr(hr), _ptr(hf), B, T, 4 * h, 4 * w, _ptr(ws), ws.numel(),
                                        _stream(self.device)), "up")
        return hr, hf

    def rescale(self, hr: torch.Tensor, T: int, seed: int = 0, offset: int = 0, eps: Optional[torch.Tensor] = None):
        """down -> 8-bit quantise -> up (models/SelfC_model.py:213-233, one pass).  Returns (lr_u8, hr_rec)."""
        _, lr_u8, lr_q = self.down(hr, T, want_out51=False)
        rec, _ = self.up(lr_q, T, eps=eps, seed=seed, offset=offset, want_hf=False)
        return lr_u8, rec

    # ---- 8-bit frames: cv2 layout [n,H,W,3], B,G,R -------------------------------------------
    def _check_img(self, x: torch.Tensor, what: str) -> torch.Tensor:
        if not x.is_cuda or x.device != self.device:
            raise RuntimeError(f"selfc_b200: {what} must live on {self.device} (got {x.device}); there is no CPU path")
        if x.dtype != torch.uint8 or x.dim() != 4 or x.shape[3] != 3:
            raise ValueError(f"{what} must be uint8 [n,H,W,3] (cv2 layout, BGR), got {x.dtype} {tuple(x.shape)}")
        return x.contiguous()

    def _img_dims(self, x: torch.Tensor, T: int, mult: int) -> Tuple[int, int, int]:
        n, hh, ww, _ = x.shape
        if T is None or T < 1 or n % T != 0:
            raise ValueError(f"batch of {n} frames is not a multiple of the clip length T={T} (GlobalVar.set_Temporal_LEN)")
        if hh % mult or ww % mult:
            raise ValueError(f"expected H,W multiples of {mult}, got {tuple(x.shape)}")
        return n // T, hh, ww

    def frames_from_u8(self, img: torch.Tensor) -> torch.Tensor:
        """read_img1 + BGR->RGB + HWC->CHW (data/util.py:103-115, LQGTVID_dataset.py:150-154): uint8 [n,H,W,3] -> fp32 [n,3,H,W]."""
        img = self._check_img(img, "img")
        n, hh, ww, _ = img.shape
        x = torch.empty((n, 3, hh, ww), dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            _lib.check(self._L.selfc_frames_from_u8(_ptr(img), _ptr(x), n, hh, ww, _stream(self.device)), "frames_from_u8")
        return x

    def frames_to_u8(self, x: torch.Tensor) -> torch.Tensor:
        """tensor2img per frame (utils/util.py:104-133): fp32 [n,3,H,W] RGB -> uint8 [n,H,W,3] BGR."""
        x = self._check_in(x, "x")
        n, c, hh, ww = x.shape
        if c != 3:
            raise ValueError(f"expected [n,3,H,W], got {tuple(x.shape)}")
        img = torch.empty((n, hh, ww, 3), dtype=torch.uint8, device=self.device)
        with torch.cuda.device(self.device):
            _lib.check(self._L.selfc_frames_to_u8(_ptr(x), _ptr(img), n, hh, ww, _stream(self.device)), "frames_to_u8")
        return img

    def down_u8(self, hr_img: torch.Tensor, T: int, want_q: bool = False):
        """8-bit HR frames -> 8-bit LR frames (+ the LR on the 1/255 fp32 grid, NCHW, when want_q)."""
        hr_img = self._check_img(hr_img, "hr_img")
        B, H, W = self._img_dims(hr_img, T, 4)
        ws = self._workspace(B, T, H // 4, W // 4)
        lr_img = torch.empty((B * T, H // 4, W // 4, 3), dtype=torch.uint8, device=self.device)
        lr_q = torch.empty((B * T, 3, H // 4, W // 4), dtype=torch.float32, device=self.device) if want_q else None
        with torch.cuda.device(self.device):
            _lib.check(self._L.selfc_down_u8(self._ctx, _ptr(hr_img), _ptr(lr_img), _ptr(lr_q), B, T, H, W, _ptr(ws), ws.numel(),
                                             _stream(self.device)), "down_u8")
        return (lr_img, lr_q) if want_q else lr_img

    def up_u8(self, lr_img: torch.Tensor, T: int, eps: Optional[torch.Tensor] = None, seed: int = 0, offset: int = 0) -> torch.Tensor:
        """8-bit LR frames (the stored LR video) -> 8-bit reconstructed HR frames."""
        lr_img = self._check_img(lr_img, "lr_img")
        B, h, w = self._img_dims(lr_img, T, 1)
        if eps is not None:
            eps = self._check_in(eps, "eps")
            if tuple(eps.shape) != (B, HF_DIM, GMM_K, T, h, w):
                raise ValueError(f"eps must be [B,48,5,T,h,w]={(B, HF_DIM, GMM_K, T, h, w)}, got {tuple(eps.shape)}")
        ws = self._workspace(B, T, h, w)
        hr_img = torch.empty((B * T, 4 * h, 4 * w, 3), dtype=torch.uint8, device=self.device)
        with torch.cuda.device(self.device):
            _lib.check(self._L.selfc_up_u8(self._ctx, _ptr(lr_img), _ptr(eps), seed & (2 ** 64 - 1), offset & (2 ** 64 - 1), _ptr(hr_img),
                                           B, T, 4 * h, 4 * w, _ptr(ws), ws.numel(), _stream(self.device)), "up_u8")
        return hr_img

    def rescale_u8(self, hr_img: torch.Tensor, T: int, seed: int = 0, offset: int = 0, eps: Optional[torch.Tensor] = None,
                   lr_out: Optional[torch.Tensor] = None, hr_out: Optional[torch.Tensor] = None):
        """8-bit frames in, 8-bit LR and reconstructed HR frames out, one call.  Returns (lr_img, hr_rec_img)."""
        hr_img = self._check_img(hr_img, "hr_img")
        B, H, W = self._img_dims(hr_img, T, 4)
        h, w = H // 4, W // 4
        if eps is not None:
            eps = self._check_in(eps, "eps")
            if tuple(eps.shape) != (B, HF_DIM, GMM_K, T, h, w):
                raise ValueError(f"eps must be [B,48,5,T,h,w]={(B, HF_DIM, GMM_K, T, h, w)}, got {tuple(eps.shape)}")
        ws = self._workspace(B, T, h, w)
        lr_img = lr_out if lr_out is not None else torch.empty((B * T, h, w, 3), dtype=torch.uint8, device=self.device)
        rec = hr_out if hr_out is not None else torch.empty((B * T, H, W, 3), dtype=torch.uint8, device=self.device)
        with torch.cuda.device(self.device):
            _lib.check(self._L.selfc_rescale_u8(self._ctx, _ptr(hr_img), _ptr(eps), seed & (2 ** 64 - 1), offset & (2 ** 64 - 1),
                                                _ptr(lr_img), _ptr(rec), B, T, H, W, _ptr(ws), ws.numel(), _stream(self.device)),
                       "rescale_u8")
        return lr_img, rec

    def rescale_host_u8(self, frames_host: torch.Tensor, lr_host: torch.Tensor, hr_host: torch.Tensor, T: int = 7,
                        seed: int = 0, offset0: int = 0) -> None:
        """rescale_host on 8-bit frames: frames_host [n,H,W,3], lr_host [n,H/4,W/4,3], hr_host [n,H,W,3], all uint8 in (pinned)
        HOST memory, cv2 layout.  A quarter of the PCIe bytes of the fp32 interface, and no fp32 frame in HBM on either side."""
        from .sharding import gop_indices
        n, H, W, _ = frames_host.shape
        dev = self.device
        if not hasattr(self, "_s_in"):
            self._s_in, self._s_out = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
        s_in, s_out = self._s_in, self._s_out
        cur = torch.cuda.current_stream(dev)
        bufs = [torch.empty((T, H, W, 3), dtype=torch.uint8, device=dev) for _ in range(2)]
        ev_free = [None, None]
        s_in.wait_stream(cur)
        for i, (ids, real) in enumerate(gop_indices(n, T)):
            buf = bufs[i & 1]
            with torch.cuda.stream(s_in):
                if ev_free[i & 1] is not None:
                    s_in.wait_event(ev_free[i & 1])
                buf[:real].copy_(frames_host[ids[0]:ids[0] + real], non_blocking=True)
                if real < T:
                    buf[real:] = buf[real - 1:real]
                ev_in = s_in.record_event()
            cur.wait_event(ev_in)
            lr_img, rec = self.rescale_u8(buf, T, seed=seed, offset=offset0 + i)
            ev_free[i & 1] = cur.record_event()
            with torch.cuda.stream(s_out):
                s_out.wait_event(ev_free[i & 1])
                lr_host[ids[0]:ids[0] + real].copy_(lr_img[:real], non_blocking=True)
                hr_host[ids[0]:ids[0] + real].copy_(rec[:real], non_blocking=True)
                lr_img.record_stream(s_out)
                rec.record_stream(s_out)
        s_out.synchronize()

    def rescale_host(self, frames_host: torch.Tensor, lr_host: torch.Tensor, hr_host: torch.Tensor, T: int = 7,
                     seed: int = 0, offset0: int = 0) -> None:
        """A whole clip held in (pinned) HOST memory -> LR codes and reconstructed HR frames in HOST memory.

        frames_host [n,3,H,W] fp32; lr_host [n,3,H/4,W/4] uint8; hr_host [n,3,H,W] fp32.  The clip is cut into GOPs of T
        frames (tail padded with copies of the last frame, models/SelfC_model.py:204-209); the H2D copy of GOP i+1 and the
        D2H copy of GOP i-1 run on side streams while GOP i computes.  Returns after everything has landed in host memory."""
        from .sharding import gop_indices
        n, _, H, W = frames_host.shape
        dev = self.device
        if not hasattr(self, "_s_in"):
            self._s_in, self._s_out = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
        s_in, s_out = self._s_in, self._s_out
        cur = torch.cuda.current_stream(dev)
        bufs = [torch.empty((T, 3, H, W), dtype=torch.float32, device=dev) for _ in range(2)]
        ev_free = [None, None]
        s_in.wait_stream(cur)
        for i, (ids, real) in enumerate(gop_indices(n, T)):
            buf = bufs[i & 1]
            with torch.cuda.stream(s_in):
                if ev_free[i & 1] is not None:
                    s_in.wait_event(ev_free[i & 1])             # the previous user of this buffer has finished reading it
                buf[:real].copy_(frames_host[ids[0]:ids[0] + real], non_blocking=True)
                if real < T:
                    buf[real:] = buf[real - 1:real]
                ev_in = s_in.record_event()
            cur.wait_event(ev_in)
            lr_u8, rec = self.rescale(buf, T, seed=seed, offset=offset0 + i)
            ev_free[i & 1] = cur.record_event()
            with torch.cuda.stream(s_out):
                s_out.wait_event(ev_free[i & 1])
                lr_host[ids[0]:ids[0] + real].copy_(lr_u8[:real], non_blocking=True)
                hr_host[ids[0]:ids[0] + real].copy_(rec[:real], non_blocking=True)
                lr_u8.record_stream(s_out)
                rec.record_stream(s_out)
        s_out.synchronize()

    # ---- per-launch timing (bench.py roofline leg) ---------------------------------------------------------------
    PROF_CLASSES = ("conv3x3", "conv5_coupling", "global_agg", "gmm_head", "sampler", "layout")

    def prof_enable(self, on: bool = True) -> None:
        _lib.check(self._L.selfc_prof_enable(self._ctx, 1 if on else 0), "prof_enable")

    def prof_read(self):
        n = len(self.PROF_CLASSES)
        ms = (C.c_double * n)()
        work = (C.c_double * n)()
        cnt = (C.c_uint64 * n)()
        _lib.check(self._L.selfc_prof_read(self._ctx, n, ms, work, cnt), "prof_read")
        return {name: {"ms": ms[i], "work": work[i], "launches": int(cnt[i])} for i, name in enumerate(self.PROF_CLASSES)}

    # ---- components (parity tests) --------------------------------------------------------------------------
    def d2dt(self, prefix: str, x: torch.Tensor, T: int) -> torch.Tensor:
        first = PARAM_INDEX[prefix + ".conv1.weight"]
        cout = {"F": 3, "G": 48, "H": 48}.get(prefix.rsplit(".", 1)[-1], 64)
        x = self._check_in(x, "x")
        B, h, w = self._clip_dims(x, T)
        ws = self._workspace(B, T, h, w)
        y = torch.empty((B * T, cout, h, w), dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            _lib.check(self._L.selfc_d2dt(self._ctx, first, _ptr(x), _ptr(y), B, T, h, w, _ptr(ws), ws.numel(),
                                          _stream(self.device)), "d2dt")
        return y

    def d2dt_backward(self, prefix: str, x: torch.Tensor, gy: torch.Tensor, T: int):
        """Backward of the dense block `prefix` (training step building block; fp32 or bf16x3 mode): x [B*T,Cin,h,w], gy [B*T,Cout,h,w]
        -> (gx, {parameter name: gradient}) with the gradients in the reference's parameter layouts."""
        first = PARAM_INDEX[prefix + ".conv1.weight"]
        x = self._check_in(x, "x")
        gy = self._check_in(gy, "gy")
        B, h, w = self._clip_dims(x, T)
        ws = self._workspace(B, T, h, w)
        gx = torch.empty_like(x)
        names = PARAM_NAMES[first:first + 10]
        grads = [torch.zeros(self._shapes[n], dtype=torch.float32, device=self.device) for n in names]
        ptrs = (C.c_void_p * 10)(*[g.data_ptr() for g in grads])
        with torch.cuda.device(self.device):
            _lib.check(self._L.selfc_d2dt_backward(self._ctx, first, _ptr(x), _ptr(gy), _ptr(gx), ptrs, B, T, h, w, _ptr(ws), ws.numel(),
                                                   _stream(self.device)), "d2dt_backward")
        return gx, dict(zip(names, grads))

    def invblock_backward(self, blk: int, rev: bool, z_in: torch.Tensor, gz: torch.Tensor, T: int):
        """Backward of InvBlockExp `operations.{blk+1}` in the forward / reverse direction (fp32 or bf16x3 mode): z_in, gz [B*T,51,h,w]
        -> (gradient w.r.t. z_in, {parameter name: gradient})."""
        z_in = self._check_in(z_in, "z_in")
        gz = self._check_in(gz, "gz").clone()
        B, h, w = self._clip_dims(z_in, T)
        ws = self._workspace(B, T, h, w)
        first = PARAM_INDEX[f"operations.{blk + 1}.F.conv1.weight"]
        names = PARAM_NAMES[first:first + 30]
        grads = [torch.zeros(self._shapes[n], dtype=torch.float32, device=self.device) for n in names]
        ptrs = (C.c_void_p * 30)(*[g.data_ptr() for g in grads])
        with torch.cuda.device(self.device):
            _lib.check(self._L.selfc_invblock_backward(self._ctx, blk, 1 if rev else 0, _ptr(z_in), _ptr(gz), ptrs, B, T, h, w, _ptr(ws),
                                                       ws.numel(), _stream(self.device)), "invblock_backward")
        return gz, dict(zip(names, grads))

    def _tape(self, B: int, T: int, h: int, w: int) -> torch.Tensor:
        n = int(self._L.selfc_train_tape_bytes(B, T, h, w))
        t = getattr(self, "_tape_buf", None)
        if t is None or t.numel() < n:
            t = torch.empty(n, dtype=torch.uint8, device=self.device)
            self._tape_buf = t
        return t

    def head_sampler_backward(self, feat: torch.Tensor, gv: torch.Tensor, T: int, eps: Optional[torch.Tensor] = None, seed: int = 0,
                              offset: int = 0):
        """Backward of tail_gmm + the soft-GMM sampler (fp32 or bf16x3 mode): feat [B*T,64,h,w], gv [B*T,48,h,w] -> (gfeat, grads)."""
        feat = self._check_in(feat, "feat")
        gv = self._check_in(gv, "gv")
        if eps is not None:
            eps = _dev_check(eps)
        B, h, w = self._clip_dims(feat, T)
        ws = self._workspace(B, T, h, w)
        tape = self._tape(B, T, h, w)
        first = PARAM_INDEX["stp_net.tail_gmm.1.weight"]
        names = PARAM_NAMES[first:first + 6]
        grads = [torch.zeros(self._shapes[n], dtype=torch.float32, device=self.device) for n in names]
        ptrs = (C.c_void_p * 6)(*[g.data_ptr() for g in grads])
        gfeat = torch.empty_like(feat)
        with torch.cuda.device(self.device):
            _lib.check(self._L.selfc_head_sampler_backward(self._ctx, _ptr(feat), _ptr(gv), _ptr(eps), seed, offset, _ptr(gfeat), ptrs,
                                                           B, T, h, w, _ptr(ws), ws.numel(), _ptr(tape), tape.numel(),
                                                           _stream(self.device)), "head_sampler_backward")
        return gfeat, dict(zip(names, grads))

    def global_agg_backward(self, prefix: str, x: torch.Tensor, gout: torch.Tensor, T: int):
        """Backward of GlobalAgg `prefix` (fp32 or bf16x3 mode): x, gout [B*T,64,h,w] -> (gx, grads of its eight parameters)."""
        x = self._check_in(x, "x")
        gout = self._check_in(gout, "gout")
        B, h, w = self._clip_dims(x, T)
        ws = self._workspace(B, T, h, w)
        tape = self._tape(B, T, h, w)
        first = PARAM_INDEX[prefix + ".fc.weight"]
        names = PARAM_NAMES[first:first + 8]
        grads = [torch.zeros(self._shapes[n], dtype=torch.float32, device=self.device) for n in names]
        ptrs = (C.c_void_p * 8)(*[g.data_ptr() for g in grads])
        gx = torch.empty_like(x)
        with torch.cuda.device(self.device):
            _lib.check(self._L.selfc_global_agg_backward(self._ctx, first, _ptr(x), _ptr(gout), _ptr(gx), ptrs, B, T, h, w, _ptr(ws),
                                                         ws.numel(), _ptr(tape), tape.numel(), _stream(self.device)), "global_agg_backward")
        return gx, dict(zip(names, grads))

    def train_grads(self, hr: torch.Tensor, ref_l: torch.Tensor, T: int, eps: Optional[torch.Tensor] = None, seed: int = 0,
                    offset: int = 0, grads: Optional[Sequence[torch.Tensor]] = None):
        """Forward + backward of one training step (SelfC_model.py:148-170; fp32 or bf16x3 mode): hr [B*T,3,H,W], ref_l [B*T,3,H/4,W/4]
        -> ({parameter name: gradient}, losses tensor [total, l_forw_fit, l_back_rec]).  `grads` (354 fp32 device tensors in
        PARAM_NAMES order) are accumulated into when given, otherwise fresh zero tensors are used."""
        hr = self._check_in(hr, "hr")
        ref_l = self._check_in(ref_l, "ref_l")
        if eps is not None:
            eps = _dev_check(eps)
        B, H, W = self._clip_dims(hr, T)
        ws = self._workspace(B, T, H // 4, W // 4)
        tape = self._tape(B, T, H // 4, W // 4)
        if grads is None:
            grads = [torch.zeros(self._shapes[n], dtype=torch.float32, device=self.device) for n in PARAM_NAMES]
        ptrs = (C.c_void_p * NUM_PARAMS)(*[g.data_ptr() for g in grads])
        losses = torch.zeros(3, dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            _lib.check(self._L.selfc_train_grads(self._ctx, _ptr(hr), _ptr(ref_l), _ptr(eps), seed, offset, ptrs, NUM_PARAMS, _ptr(losses),
                                                 B, T, H, W, _ptr(ws), ws.numel(), _ptr(tape), tape.numel(), _stream(self.device)),
                       "train_grads")
        return dict(zip(PARAM_NAMES, grads)), losses

    def conv3x3(self, prefix: str, k: int, x: torch.Tensor, T: int) -> torch.Tensor:
        """conv{k+1} of the dense block `prefix` on its concatenated input x [B*T,Cin+32k,h,w] -> [B*T,32,h,w]."""
        first = PARAM_INDEX[prefix + ".conv1.weight"]
        x = self._check_in(x, "x")
        B, h, w = self._clip_dims(x, T)
        ws = self._workspace(B, T, h, w)
        y = torch.empty((B * T, 32, h, w), dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            _lib.check(self._L.selfc_conv3x3(self._ctx, first, k, _ptr(x), _ptr(y), B, T, h, w, _ptr(ws), ws.numel(),
                                             _stream(self.device)), "conv3x3")
        return y

    def global_agg(self, prefix: str, x: torch.Tensor, T: int):
        first = PARAM_INDEX[prefix + ".fc.weight"]
        x = self._check_in(x, "x")
        B, h, w = self._clip_dims(x, T)
        ws = self._workspace(B, T, h, w)
        y = torch.empty_like(x)
        wmat = torch.empty((B, T, T), dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            _lib.check(self._L.selfc_global_agg(self._ctx, first, _ptr(x), _ptr(y), _ptr(wmat), B, T, h, w, _ptr(ws),
                                                ws.numel(), _stream(self.device)), "global_agg")
        return y, wmat


# ---- context-free components -----------------------------------------------------------------------------------
def _dev_check(x: torch.Tensor) -> torch.Tensor:
    if not x.is_cuda:
        raise RuntimeError("selfc_b200 runs on CUDA (sm_100a) only; there is no CPU fallback")
    return x.to(torch.float32).contiguous()


def fa_forward(x: torch.Tensor, k: int = 4) -> torch.Tensor:
    """FrequencyAnalyzer.forward(rev=False): k=4 the rescaler's (SelfC_GMM_arch_inv.py:62-78), k=2 the codec model's
    (SelfC_Codec_arch_inv.py:78-94)."""
    x = _dev_check(x)
    n, c, hh, ww = x.shape
    if k == 2:
        if c != 3 or hh % 2 or ww % 2:
            raise ValueError(f"fa_forward(k=2) expects [N,3,H,W] with even H,W, got {tuple(x.shape)}")
        out = torch.empty((n, 15, hh // 2, ww // 2), dtype=torch.float32, device=x.device)
        with torch.cuda.device(x.device):
            _lib.check(_lib.lib().selfc_fa2_fwd(_ptr(x), _ptr(out), n, hh, ww, _stream(x.device)), "fa2_fwd")
        return out
    if k != 4:
        raise ValueError("FrequencyAnalyzer is built for k=4 and k=2")
    out = torch.empty((n, 51, hh // 4, ww // 4), dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        _lib.check(_lib.lib().selfc_fa_fwd(_ptr(x), _ptr(out), n, hh, ww, _stream(x.device)), "fa_fwd")
    return out


def fa_reverse(z: torch.Tensor, k: int = 4) -> torch.Tensor:
    z = _dev_check(z)
    n, c, h, w = z.shape
    if k == 2:
        if c != 15:
            raise ValueError("fa_reverse(k=2) expects 15 channels")
        y = torch.empty((n, 3, 2 * h, 2 * w), dtype=torch.float32, device=z.device)
        with torch.cuda.device(z.device):
            _lib.check(_lib.lib().selfc_fa2_rev(_ptr(z), _ptr(y), n, h, w, _stream(z.device)), "fa2_rev")
        return y
    if c != 51:
        raise ValueError("fa_reverse expects 51 channels")
    y = torch.empty((n, 3, 4 * h, 4 * w), dtype=torch.float32, device=z.device)
    with torch.cuda.device(z.device):
        _lib.check(_lib.lib().selfc_fa_rev(_ptr(z), _ptr(y), n, h, w, _stream(z.device)), "fa_rev")
    return y


def haar_forward(x: torch.Tensor) -> torch.Tensor:
    """HaarDownsampling.forward(rev=False) (SelfC_arch_inv.py:66-74): [N,C,H,W] -> [N,4C,H/2,W/2]."""
    x = _dev_check(x)
    n, c, hh, ww = x.shape
    if hh % 2 or ww % 2:
        raise ValueError(f"haar_forward expects even H,W, got {tuple(x.shape)}")
    out = torch.empty((n, 4 * c, hh // 2, ww // 2), dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        _lib.check(_lib.lib().selfc_haar_fwd(_ptr(x), _ptr(out), n, c, hh, ww, _stream(x.device)), "haar_fwd")
    return out


def haar_reverse(z: torch.Tensor) -> torch.Tensor:
    """HaarDownsampling.forward(rev=True) (:75-82): [N,4C,h,w] -> [N,C,2h,2w]."""
    z = _dev_check(z)
    n, c4, h, w = z.shape
    if c4 % 4:
        raise ValueError("haar_reverse expects a multiple of 4 channels")
    y = torch.empty((n, c4 // 4, 2 * h, 2 * w), dtype=torch.float32, device=z.device)
    with torch.cuda.device(z.device):
        _lib.check(_lib.lib().selfc_haar_rev(_ptr(z), _ptr(y), n, c4 // 4, h, w, _stream(z.device)), "haar_rev")
    return y


def quantize(x: torch.Tensor):
    x = _dev_check(x)
    q8 = torch.empty(x.shape, dtype=torch.uint8, device=x.device)
    qf = torch.empty_like(x)
    with torch.cuda.device(x.device):
        _lib.check(_lib.lib().selfc_quantize(_ptr(x), _ptr(q8), _ptr(qf), x.numel(), _stream(x.device)), "quantize")
    return q8, qf


_GAUSS13 = {}


def gaussian_kernel_13(sigma: float = 1.6) -> torch.Tensor:
    """The 13x13 taps the reference builds with scipy.ndimage.gaussian_filter on a dirac (models/Guassian.py:16-22):
    a separable, 4-sigma-truncated, normalised Gaussian (truncate=4 -> radius 6 at sigma 1.6)."""
    import math
    g = [math.exp(-0.5 * (i / sigma) ** 2) for i in range(-6, 7)]
    ssum = sum(g)
    g = torch.tensor([v / ssum for v in g], dtype=torch.float64)
    return torch.outer(g, g).to(torch.float32)


def gaussian_downsample(x: torch.Tensor) -> torch.Tensor:
    """LR_ref for `distortion: sr_bd`: [N,C,H,W] -> [N,C,H/4,W/4] (models/Guassian.py:7-52)."""
    x = _dev_check(x)
    n, c, hh, ww = x.shape
    key = x.device
    if key not in _GAUSS13:
        _GAUSS13[key] = gaussian_kernel_13().to(x.device).contiguous()
    y = torch.empty((n, c, hh // 4, ww // 4), dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        _lib.check(_lib.lib().selfc_gaussian_down(_ptr(x), _ptr(_GAUSS13[key]), _ptr(y), n, c, hh, ww, _stream(x.device)),
                   "gaussian_down")
    return y


def gmm_sample(params: torch.Tensor, T: int, eps: Optional[torch.Tensor] = None, seed: int = 0, offset: int = 0):
    params = _dev_check(params)
    bt, c, h, w = params.shape
    if c != 720 or bt % T:
        raise ValueError("gmm_sample expects [B*T,720,h,w]")
    if eps is not None:
        eps = _dev_check(eps)
    v = torch.empty((bt, HF_DIM, h, w), dtype=torch.float32, device=params.device)
    with torch.cuda.device(params.device):
        _lib.check(_lib.lib().selfc_gmm_sample(_ptr(params), _ptr(eps), seed, offset, _ptr(v), bt // T, T, h, w,
                                               _stream(params.device)), "gmm_sample")
    return v


def gmm_params_to_planar(params: torch.Tensor) -> torch.Tensor:
    """[B*T,720,h,w] in the reference's channel order hf*15 + k*3 + j (SelfC_GMM_arch_inv.py:383-388) -> the planar quads the
    tcgen05 head writes: [180][M][4] with quad = j*60 + k*12 + i holding hf = 4i..4i+3 (j: 0 logit, 1 log-sigma, 2 mu)."""
    bt, c, h, w = params.shape
    m = bt * h * w
    # [M, i, e, k, j] -> [j, k, i, M, e]
    return params.permute(0, 2, 3, 1).reshape(m, HF_DIM // 4, 4, GMM_K, 3).permute(4, 3, 1, 0, 2).reshape(180, m, 4).contiguous()


def gmm_latent_from_planar(z: torch.Tensor, bt: int, h: int, w: int) -> torch.Tensor:
    """planar latent state [13][M][4] (quad 0 = LR frame, quads 1..12 = HF) -> v [B*T,48,h,w]"""
    return z[1:].permute(1, 0, 2).reshape(bt, h, w, HF_DIM).permute(0, 3, 1, 2).contiguous()


def gmm_sample_planar(params: torch.Tensor, T: int, eps: Optional[torch.Tensor] = None, seed: int = 0, offset: int = 0,
                      form: int = -1):
    """The sampler the bf16 mode launches, driven from reference-layout tensors: params [B*T,720,h,w] (channel hf*15+k*3+j,
    SelfC_GMM_arch_inv.py:383-388) are permuted into the planar quads the tcgen05 head writes, the planar latent comes back
    as v [B*T,48,h,w].  form 0 / 1 selects the thread-per-pixel / warp-split kernel, -1 the default."""
    params = _dev_check(params)
    bt, c, h, w = params.shape
    if c != 720 or bt % T:
        raise ValueError("gmm_sample_planar expects [B*T,720,h,w]")
    if eps is not None:
        eps = _dev_check(eps)
    m = bt * h * w
    planar = gmm_params_to_planar(params)
    z = torch.zeros((1 + HF_DIM // 4, m, 4), dtype=torch.float32, device=params.device)
    with torch.cuda.device(params.device):
        _lib.check(_lib.lib().selfc_gmm_sample_planar(_ptr(planar), _ptr(eps), seed, offset, _ptr(z), bt // T, T, h, w, form,
                                                      _stream(params.device)), "gmm_sample_planar")
    return gmm_latent_from_planar(z, bt, h, w)


def export_eps(B: int, T: int, h: int, w: int, seed: int, offset: int, device) -> torch.Tensor:
    eps = torch.empty((B, HF_DIM, GMM_K, T, h, w), dtype=torch.float32, device=device)
    with torch.cuda.device(eps.device):
        _lib.check(_lib.lib().selfc_export_eps(_ptr(eps), seed, offset, B, T, h, w, _stream(eps.device)), "export_eps")
    return eps


def launch_count() -> int:
    return int(_lib.lib().selfc_launch_count())
