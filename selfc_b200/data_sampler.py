"""Iteration-oriented distributed sampler with the reference's semantics (codes/data/data_sampler.py:12-66).

`DistIterSampler(dataset, num_replicas, rank, ratio)` enlarges an epoch to `ratio` passes over the dataset so the
DataLoader is restarted rarely (train.py:148-156 uses ratio 200), shuffles the enlarged index range with a generator
seeded by the epoch, folds the indices back onto the dataset (`% len(dataset)`) and gives rank r every
`num_replicas`-th index starting at r.  Same constructor, `set_epoch`, `__len__` and -- for the same torch version --
the same index stream as the reference class (pinned by tests/golden/sampler.npz, produced by that class).
"""
from __future__ import annotations

import math

import torch
import torch.distributed as dist
from torch.utils.data.sampler import Sampler


class DistIterSampler(Sampler):
    def __init__(self, dataset, num_replicas=None, rank=None, ratio=100):
        if num_replicas is None:
            if not (dist.is_available() and dist.is_initialized()):
                raise RuntimeError("Requires distributed package to be available")
            num_replicas = dist.get_world_size()
        if rank is None:
            if not (dist.is_available() and dist.is_initialized()):
                raise RuntimeError("Requires distributed package to be available")
            rank = dist.get_rank()
        self.dataset = dataset
        self.num_replicas = num_replicas
        self.rank = rank
        self.epoch = 0
        self.num_samples = int(math.ceil(len(self.dataset) * ratio / self.num_replicas))
        self.total_size = self.num_samples * self.num_replicas

    def __iter__(self):
        g = torch.Generator()
        g.manual_seed(self.epoch)                       # deterministic shuffle per epoch, identical on every rank
        order = torch.randperm(self.total_size, generator=g)
        order = order % len(self.dataset)
        mine = order[self.rank:self.total_size:self.num_replicas].tolist()
        assert len(mine) == self.num_samples
        return iter(mine)

    def __len__(self):
        return self.num_samples

    def set_epoch(self, epoch):
        self.epoch = epoch
