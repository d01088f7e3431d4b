"""Training step of the rescaling path: the host side of `SelfCModel.optimize_parameters`
(models/SelfC_model.py:148-183) on top of the C-ABI -- `selfc_train_grads` (forward + backward, fp32-FMA kernels) and
`selfc_adam_step` (clip_grad_norm_ + Adam on a flat gradient buffer).

Data parallelism (train.py:94-100 of the reference wraps netG in DistributedDataParallel): one process per GPU, every rank runs
the step on its own clips, ONE `torch.distributed.all_reduce` (NCCL over NVLink) of the flat 3,365,038-element gradient, the
1/world average folded into the optimiser kernel -- clipping therefore sees the averaged gradient, as under DDP.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, Optional, Sequence

import torch

from . import _lib
from .engine import NUM_PARAMS, PARAM_NAMES, Engine, _ptr, _stream


def multistep_lr(base_lr: float, step: int, milestones: Sequence[int], gamma: float) -> float:
    """lr_scheduler.MultiStepLR_Restart without restarts (train_rescaling_selfc_large.yml: lr_steps, lr_gamma)."""
    return base_lr * (gamma ** sum(1 for m in (milestones or []) if step >= m))


def all_reduce_sum(flat: torch.Tensor) -> float:
    """SUM all-reduce of a flat gradient buffer over the default process group (NCCL on the GPU box, gloo in the CPU tests);
    returns the scale (1/world) that turns the sum into the data-parallel mean.  No-op without a process group."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(flat, op=dist.ReduceOp.SUM)
        return 1.0 / dist.get_world_size()
    return 1.0


def broadcast_parameters(params: Sequence[torch.Tensor], src: int = 0) -> bool:
    """DistributedDataParallel broadcasts rank `src`'s parameters when it wraps the module (train.py:94-100 of the reference);
    the flat-gradient all-reduce only keeps the replicas identical if they START identical, so the trainer does the same.
    Returns True when a broadcast happened (a process group with more than one rank exists)."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1):
        return False
    with torch.no_grad():
        for p in params:
            dist.broadcast(p.data, src=src)
    return True


def shard_batch(n_clips: int, world: int, rank: int) -> range:
    """Per-rank slice of a global batch of clips (data/__init__.py:13-14: batch_size // world per rank)."""
    per = n_clips // world
    return range(rank * per, (rank + 1) * per)


class Trainer:
    """Flat gradient / Adam-moment buffers over the network's parameters (which stay separate tensors in the reference
    layout, so `state_dict()` is unchanged) and the step itself."""

    def __init__(self, net, device: torch.device, lr: float = 1e-4, betas=(0.9, 0.999), eps: float = 1e-8, weight_decay: float = 0.0,
                 max_norm: Optional[float] = 10.0):
        if net.precision not in ("fp32", "bf16x3", "fp32_tc", None):
            raise RuntimeError("the training step runs in fp32 mode (fp32-FMA kernels) or in bf16x3 mode (the same arithmetic to 16 "
                               "mantissa bits per operand on the tensor cores); plain bf16 has no backward")
        self.net, self.device = net, device
        self.engine: Engine = net._engine_for(device)
        own = dict(net.named_parameters())
        self.params: List[torch.Tensor] = [own[n] for n in PARAM_NAMES]
        sizes = [p.numel() for p in self.params]
        offs = [0]
        for n in sizes:
            offs.append(offs[-1] + n)
        self.total = offs[-1]
        self.flat_grad = torch.zeros(self.total, dtype=torch.float32, device=device)
        self.grad_views = [self.flat_grad[offs[i]:offs[i + 1]].view(p.shape) for i, p in enumerate(self.params)]
        self.m = torch.zeros_like(self.flat_grad)
        self.v = torch.zeros_like(self.flat_grad)
        self._offsets = torch.tensor(offs, dtype=torch.int64, device=device)
        self._ptrs = torch.tensor([p.data_ptr() for p in self.params], dtype=torch.int64, device=device)
        self._sq = torch.zeros(1, dtype=torch.float32, device=device)
        self.lr, self.betas, self.eps, self.weight_decay, self.max_norm = lr, betas, eps, weight_decay, max_norm
        self.step_count = 0
        self._ptr_key = None
        self.broadcast_parameters()

    def broadcast_parameters(self, src: int = 0) -> None:
        if broadcast_parameters(self.params, src):
            self._bump_versions()

    def _bump_versions(self) -> None:
        """The optimiser kernel (and the broadcast above) write the parameters behind autograd's back: bump the version counters
        so that EVERY engine caching packed weights (keyed on (data_ptr, _version)) re-packs, not only this trainer's."""
        with torch.no_grad():
            torch._foreach_add_(self.params, 0.0)       # a handful of multi-tensor launches instead of 354 (0.6 ms of a 44 ms step)

    def _refresh_ptrs(self):
        cur = [p.data_ptr() for p in self.params]
        if cur != self._ptr_key:
            # selfc_adam_step writes through raw pointers: every parameter must be a contiguous fp32 tensor on this device
            for name, p in zip(PARAM_NAMES, self.params):
                same_dev = p.device.type == "cuda" and (self.device.index is None or p.device.index == self.device.index)
                if p.dtype != torch.float32 or not p.is_contiguous() or not same_dev:
                    raise RuntimeError(f"parameter {name} must be a contiguous float32 tensor on {self.device} for the optimiser "
                                       f"kernel (got {p.dtype}, contiguous={p.is_contiguous()}, {p.device})")
            self._ptrs = torch.tensor(cur, dtype=torch.int64, device=self.device)
            self._ptr_key = cur

    def grads_and_losses(self, real_h: torch.Tensor, ref_l: torch.Tensor, t: int, eps: Optional[torch.Tensor] = None, seed: int = 0,
                         offset: int = 0):
        """zero_grad + forward + backward; returns the losses tensor [total, l_forw_fit, l_back_rec] (device)."""
        self.engine = self.net._engine_for(self.device)          # re-packs the kernels' weight images if parameters changed
        self.flat_grad.zero_()
        _, losses = self.engine.train_grads(real_h, ref_l, t, eps=eps, seed=seed, offset=offset, grads=self.grad_views)
        return losses

    def all_reduce(self) -> float:
        """SUM all-reduce of the flat gradient over the data-parallel group; returns the 1/world scale for the optimiser."""
        return all_reduce_sum(self.flat_grad)

    def apply(self, gscale: float = 1.0, lr: Optional[float] = None):
        """clip_grad_norm_(max_norm) + Adam step, in place on the parameters."""
        self.step_count += 1
        self._refresh_ptrs()
        L = _lib.lib()
        with torch.cuda.device(self.device):
            _lib.check(L.selfc_adam_step(_ptr(self._ptrs), _ptr(self._offsets), NUM_PARAMS, self.total, _ptr(self.flat_grad), _ptr(self.m),
                                         _ptr(self.v), _ptr(self._sq), float(gscale), float(self.max_norm or 0.0),
                                         float(self.lr if lr is None else lr), float(self.betas[0]), float(self.betas[1]), float(self.eps),
                                         float(self.weight_decay), int(self.step_count), _stream(self.device)), "adam_step")
        self._bump_versions()          # parameters were written through raw pointers: every cached engine re-packs
        self.engine.invalidate()

    def step(self, real_h, ref_l, t, eps=None, seed=0, offset=0, lr=None):
        losses = self.grads_and_losses(real_h, ref_l, t, eps=eps, seed=seed, offset=offset)
        self.apply(self.all_reduce(), lr=lr)
        return losses

    def named_grads(self) -> Dict[str, torch.Tensor]:
        return dict(zip(PARAM_NAMES, self.grad_views))
