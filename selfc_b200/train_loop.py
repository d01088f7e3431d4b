"""Training driver with the reference's control flow (codes/train.py:88-330) around the CUDA training step.

    python -m selfc_b200.train_loop -opt selfc_b200/configs/selfc_large_train_synthetic.yml [--launcher pytorch]

What is mirrored: option parsing, distributed init from the launcher's environment (train.py:19-27), resume-state
loading, the seed, `DistIterSampler` with the x200 epoch enlargement, the epoch/iteration loop
(`feed_data -> optimize_parameters -> update_learning_rate`), the log line format, checkpoint + training-state files
every `save_checkpoint_freq` iterations.  What differs, on purpose:
  * one process per GPU; the gradient all-reduce inside `optimize_parameters` replaces DistributedDataParallel;
  * resume really restores the Adam moments, step count and learning rate (the reference's `resume_training` body is
    commented out, base_model.py:122-133, so it silently restarts the optimiser);
  * validation (`cal_metric`, train.py:28-86) is `validate()` below on whatever loaders the caller passes; PNG writing is
    left to the caller (`Engine.frames_to_u8` gives the `tensor2img` bytes).
A `SyntheticClips` dataset stands in for Vimeo-90k septuplets when the YAML has no `dataroot_GT` (this image has no data).
"""
from __future__ import annotations

import argparse
import logging
import math
import os
import random
from typing import Dict, Iterable, Optional

import numpy as np
import torch
import torch.distributed as dist

from . import options as option
from .data_sampler import DistIterSampler
from .global_var import GlobalVar
from .model import create_model

DATASET_RATIO = 200          # train.py:143


class SyntheticClips(torch.utils.data.Dataset):
    """`n` deterministic smooth clips shaped like the training crops: item -> {'GT': [3,T,S,S] in [0,1]} (the dict
    LQGTVIDDataset.__getitem__ returns, LQGTVID_dataset.py:157-229, without the file paths)."""

    def __init__(self, n: int = 64, t: int = 7, size: int = 144, seed: int = 0):
        self.n, self.t, self.size, self.seed = n, t, size, seed

    def __len__(self):
        return self.n

    def __getitem__(self, i):
        g = torch.Generator().manual_seed(self.seed * 1000003 + int(i))
        base = torch.rand(1, 3 * self.t, self.size // 8, self.size // 8, generator=g)
        x = torch.nn.functional.interpolate(base, size=(self.size, self.size), mode="bicubic", align_corners=False)
        x = (x + 0.02 * torch.randn(x.shape, generator=g)).clamp(0, 1)
        x = torch.round(x * 255.0) / 255.0
        return {"GT": x.reshape(3, self.t, self.size, self.size), "GT_path": f"synthetic/{i:05d}"}


def init_dist(backend: str = "nccl"):
    """train.py:19-27: rank / world size from the launcher's environment, one GPU per process."""
    rank = int(os.environ["RANK"])
    if torch.cuda.is_available():
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank % max(1, torch.cuda.device_count()))))
    if not dist.is_initialized():
        dist.init_process_group(backend=backend)
    return dist.get_rank(), dist.get_world_size()


def set_random_seed(seed: int):
    """utils/util.py:77-81."""
    random.seed(seed)
    np.random.seed(seed)
    torch.manual_seed(seed)
    torch.cuda.manual_seed_all(seed)


def epochs_needed(n_items: int, batch_size: int, niter: int, distributed: bool) -> int:
    """train.py:146-152: iterations per epoch from the dataset size, x200 under the distributed sampler."""
    train_size = int(math.ceil(n_items / batch_size))
    if distributed:
        return int(math.ceil(niter / (train_size * DATASET_RATIO)))
    return int(math.ceil(niter / train_size))


def log_message(epoch: int, step: int, lr: float, logs: Dict[str, float]) -> str:
    """The reference's log line (train.py:255-259)."""
    msg = "<epoch:{:3d}, iter:{:8,d}, lr:{:.3e}> ".format(epoch, step, lr)
    for k, v in logs.items():
        msg += "{:s}: {:.4e} ".format(k, v)
    return msg


def train(opt, model, train_loader: Iterable, train_sampler: Optional[DistIterSampler] = None, resume_state: Optional[dict] = None,
          total_epochs: Optional[int] = None, rank: int = -1, logger: Optional[logging.Logger] = None) -> int:
    """The loop of train.py:237-330 (training half).  Returns the last completed iteration."""
    logger = logger or logging.getLogger("base")
    total_iters = int(opt["train"]["niter"])
    if resume_state:
        logger.info("Resuming training from epoch: {}, iter: {}.".format(resume_state["epoch"], resume_state["iter"]))
        start_epoch, current_step = int(resume_state["epoch"]), int(resume_state["iter"])
        model.resume_training(resume_state)
    else:
        start_epoch, current_step = 0, 0
    if total_epochs is None:
        total_epochs = 1 << 30
    logger.info("Start training from epoch: {:d}, iter: {:d}".format(start_epoch, current_step))
    print_freq = int(opt["logger"]["print_freq"]) if opt["logger"] and opt["logger"]["print_freq"] else 0
    save_freq = int(opt["logger"]["save_checkpoint_freq"]) if opt["logger"] and opt["logger"]["save_checkpoint_freq"] else 0
    for epoch in range(start_epoch, total_epochs + 1):
        if train_sampler is not None:
            train_sampler.set_epoch(epoch)
        for train_data in train_loader:
            current_step += 1
            if current_step > total_iters:
                return current_step - 1
            model.feed_data(train_data)
            model.optimize_parameters(current_step)
            model.update_learning_rate(current_step, warmup_iter=opt["train"]["warmup_iter"])
            if print_freq and current_step % print_freq == 0 and rank <= 0:
                logger.info(log_message(epoch, current_step, model.get_current_learning_rate(), model.get_current_log()))
            if save_freq and current_step % save_freq == 0 and rank <= 0:
                logger.info("Saving models and training states.")
                model.save(current_step)
                model.save_training_state(epoch, current_step)
        if current_step >= total_iters:
            break
    return min(current_step, total_iters)


def validate(model, val_loader: Iterable) -> Dict[str, float]:
    """cal_metric (train.py:28-86) without the PNG dumps: per-clip RGB / Y PSNR + SSIM of SR vs GT and LR vs LR_ref, averaged."""
    from .metrics import clip_metrics
    acc: Dict[str, list] = {}
    for val_data in val_loader:
        model.feed_data(val_data)
        model.test()
        vis = model.get_current_visuals()
        for k, v in clip_metrics(vis["SR"], vis["GT"], vis["LR"], vis["LR_ref"]).items():
            acc.setdefault(k, []).extend(v)
    return {k: float(sum(v) / len(v)) for k, v in acc.items() if v}


def main(argv=None):
    p = argparse.ArgumentParser()
    p.add_argument("-opt", type=str, required=True, help="Path to option YAML file.")
    p.add_argument("--launcher", choices=["none", "pytorch"], default="none", help="job launcher")
    p.add_argument("--local_rank", type=int, default=0)
    p.add_argument("--niter", type=int, default=None, help="override train.niter (smoke runs)")
    args = p.parse_args(argv)
    opt = option.parse(args.opt, is_train=True)
    if args.launcher == "none":
        opt["dist"], rank, world = False, -1, 1
    else:
        opt["dist"] = True
        rank, world = init_dist()
    resume_state = None
    if opt["path"].get("resume_state"):
        resume_state = torch.load(opt["path"]["resume_state"], map_location="cpu", weights_only=False)
        option.check_resume(opt, resume_state["iter"])          # weights = <models>/<iter>_G.pth (train.py:122)
    if rank <= 0:
        for key in ("models", "training_state"):
            os.makedirs(opt["path"][key], exist_ok=True)
    logging.basicConfig(level=logging.INFO, format="%(asctime)s %(message)s")
    logger = logging.getLogger("base")
    opt = option.dict_to_nonedict(opt)
    if args.niter:
        opt["train"]["niter"] = args.niter
    seed = opt["train"]["manual_seed"]
    if seed is None:
        seed = random.randint(1, 10000)
        if opt["dist"]:
            # every rank must build the same initial weights and draw the same synthetic clips: rank 0's draw wins
            box = [seed]
            dist.broadcast_object_list(box, src=0)
            seed = box[0]
    set_random_seed(int(seed))
    ds_opt = opt["datasets"]["train"]
    t = int(ds_opt["video_len"] or 7)
    GlobalVar.set_Temporal_LEN(t)
    if ds_opt["dataroot_GT"]:
        raise NotImplementedError("file-backed datasets are not part of this package: pass your own loader to train()")
    train_set = SyntheticClips(n=64, t=t, size=int(ds_opt["GT_size"] or 144), seed=int(seed))
    batch = int(ds_opt["batch_size"])
    sampler = DistIterSampler(train_set, world, rank, DATASET_RATIO) if opt["dist"] else None
    per_rank = max(1, batch // world) if opt["dist"] else batch          # data/__init__.py:13-16
    loader = torch.utils.data.DataLoader(train_set, batch_size=per_rank, shuffle=sampler is None, sampler=sampler, num_workers=0,
                                         drop_last=True, pin_memory=True)
    total_epochs = epochs_needed(len(train_set), batch, int(opt["train"]["niter"]), bool(opt["dist"]))
    model = create_model(opt)
    last = train(opt, model, loader, sampler, resume_state, total_epochs, rank, logger)
    if rank <= 0:
        logger.info("End of training at iter {:d}: {}".format(last, log_message(0, last, model.get_current_learning_rate(),
                                                                                model.get_current_log())))
    if opt["dist"]:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
