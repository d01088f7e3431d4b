"""Host logic of the multi-GPU path, on CPU: GOP decomposition, contiguous partitioning, and the metric gather under a
world_size-2 gloo process group (the N>1 path has no data-path collective)."""
import os

import torch
import torch.multiprocessing as mp

from selfc_b200 import sharding


def test_gop_indices_pad_like_the_reference():
    g = sharding.gop_indices(100)
    assert len(g) == 15 and all(len(ids) == 7 for ids, _ in g)
    assert [r for _, r in g] == [7] * 14 + [2]
    assert g[-1][0] == [98, 99, 99, 99, 99, 99, 99]
    assert sharding.gop_indices(7) == [([0, 1, 2, 3, 4, 5, 6], 7)]
    assert sharding.gop_indices(3)[0] == ([0, 1, 2, 2, 2, 2, 2], 3)


def test_partition_covers_every_unit_once():
    for n in (0, 1, 7, 15, 120, 121):
        for world in (1, 2, 3, 4, 8):
            parts = [sharding.partition(n, world, r) for r in range(world)]
            flat = [i for p in parts for i in p]
            assert flat == list(range(n))
            assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1


def _worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    units = sharding.partition(15, world, rank)                 # 15 GOPs of a 100-frame group
    local = torch.tensor([[float(u), 2.0 * u, 10.0 + u, 0.5 * u] for u in units])   # 4 metrics per unit
    allm = sharding.gather_metrics(local)
    q.put((rank, allm.tolist(), list(units)))
    dist.barrier()
    dist.destroy_process_group()


def test_metric_gather_world2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29000 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    expect = [[float(u), 2.0 * u, 10.0 + u, 0.5 * u] for u in range(15)]
    for rank, allm, units in res:
        assert allm == expect                                   # every rank sees all rows, in unit order
    assert sorted(u for _, _, units in res for u in units) == list(range(15))


# ---- training step (row a13): the gradient all-reduce and the per-rank batch split, world_size 2, gloo ------------------------
def _grad_worker(rank, world, port, q):
    import torch.distributed as dist
    from selfc_b200 import train
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    flat = torch.arange(1000, dtype=torch.float32) * (rank + 1)          # this rank's flat gradient
    scale = train.all_reduce_sum(flat)
    q.put((rank, scale, (flat * scale).tolist()[:5], list(train.shard_batch(8, world, rank))))
    dist.barrier()
    dist.destroy_process_group()


def test_gradient_allreduce_world2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31000 + os.getpid() % 2000
    procs = [ctx.Process(target=_grad_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, scale, head, clips in res:
        assert scale == 0.5
        assert head == [0.0, 1.5, 3.0, 4.5, 6.0]                        # mean of g and 2g
        assert clips == list(range(4 * rank, 4 * rank + 4))


def _bcast_worker(rank, world, port, q):
    import torch.distributed as dist
    from selfc_b200 import train
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(100 + rank)                                        # every rank starts from DIFFERENT weights
    params = [torch.nn.Parameter(torch.randn(3, 5)), torch.nn.Parameter(torch.randn(7))]
    before = [p.detach().clone() for p in params]
    did = train.broadcast_parameters(params, src=0)
    q.put((rank, did, [p.detach().tolist() for p in params], [b.tolist() for b in before]))
    dist.barrier()
    dist.destroy_process_group()


def test_initial_parameter_broadcast_world2_gloo():
    """DDP's construction-time broadcast (train.py:94-100): after Trainer.broadcast_parameters every rank holds rank 0's weights."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 33000 + os.getpid() % 2000
    procs = [ctx.Process(target=_bcast_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, did0, after0, before0), (r1, did1, after1, before1) = res
    assert did0 and did1
    assert before0 != before1                                            # they really started apart
    assert after0 == before0 and after1 == before0                       # and both ended on rank 0's values
    from selfc_b200 import train
    assert train.broadcast_parameters([torch.nn.Parameter(torch.zeros(2))]) is False      # no process group: nothing to do


def test_check_resume_points_at_saved_weights(tmp_path):
    """options.check_resume (reference options.py:105-120, train.py:122): resuming loads <models>/<iter>_G.pth, warns when a
    pretrain path was configured, and fails when the weights that belong to the training state are missing."""
    import pytest
    from selfc_b200 import options
    models = tmp_path / "models"
    models.mkdir()
    opt = {"path": {"resume_state": str(tmp_path / "training_state" / "200.state"), "models": str(models), "pretrain_model_G": "/somewhere/else.pth"}}
    with pytest.raises(FileNotFoundError):
        options.check_resume(opt, 200)
    (models / "200_G.pth").write_bytes(b"x")
    options.check_resume(opt, 200)
    assert opt["path"]["pretrain_model_G"] == str(models / "200_G.pth")
    opt2 = {"path": {"resume_state": None, "models": str(models), "pretrain_model_G": "keep.pth"}}
    options.check_resume(opt2, 200)
    assert opt2["path"]["pretrain_model_G"] == "keep.pth"               # not resuming: untouched


def test_multistep_lr_and_single_process_allreduce():
    from selfc_b200 import train
    assert train.multistep_lr(1e-4, 0, [100000, 200000, 300000], 0.5) == 1e-4
    assert train.multistep_lr(1e-4, 100000, [100000, 200000, 300000], 0.5) == 5e-5
    assert train.multistep_lr(1e-4, 350000, [100000, 200000, 300000], 0.5) == 1.25e-5
    g = torch.ones(4)
    assert train.all_reduce_sum(g) == 1.0 and torch.equal(g, torch.ones(4))


# ------------------------------------------------------------------------------------------------ f4: training driver (host logic)
def test_dist_iter_sampler_matches_reference(golden_dir):
    """Index streams identical to the reference's DistIterSampler (fixture from data/data_sampler.py) when the torch version that
    generated the fixture is the one running; always: every rank gets num_samples indices and the ranks partition the epoch."""
    import os
    import numpy as np
    import torch
    from selfc_b200.data_sampler import DistIterSampler
    g = np.load(os.path.join(golden_dir, "sampler.npz"))
    same_torch = str(g["torch_version"]) == torch.__version__
    for size, world, ratio in ((37, 3, 4), (64, 2, 200), (5, 4, 1)):
        ds = list(range(size))
        for epoch in (0, 3):
            streams = []
            for rank in range(world):
                smp = DistIterSampler(ds, world, rank, ratio)
                smp.set_epoch(epoch)
                idx = list(iter(smp))
                assert len(idx) == len(smp) == int(g[f"s{size}_w{world}_r{ratio}_k{rank}_len"])
                if same_torch:
                    assert idx == g[f"s{size}_w{world}_r{ratio}_k{rank}_e{epoch}"].tolist()
                streams.append(idx)
            # interleaving the ranks gives the (seeded) permutation of the enlarged epoch folded onto the dataset
            gen = torch.Generator().manual_seed(epoch)
            total = len(streams[0]) * world
            full = (torch.randperm(total, generator=gen) % size).tolist()
            inter = [streams[i % world][i // world] for i in range(total)]
            assert inter == full
    with __import__("pytest").raises(RuntimeError):
        DistIterSampler(list(range(4)))          # no process group, no explicit world/rank


def test_training_driver_helpers():
    from selfc_b200 import train_loop
    # train.py:146-152
    assert train_loop.epochs_needed(64, 16, 400000, distributed=False) == 100000
    assert train_loop.epochs_needed(64, 16, 400000, distributed=True) == 500
    msg = train_loop.log_message(3, 12000, 5e-5, {"l_forw_fit": 1.25e-3, "loss": 0.5})
    assert msg == "<epoch:  3, iter:  12,000, lr:5.000e-05> l_forw_fit: 1.2500e-03 loss: 5.0000e-01 "
    ds = train_loop.SyntheticClips(n=4, t=3, size=32, seed=1)
    a, b = ds[2]["GT"], ds[2]["GT"]
    assert a.shape == (3, 3, 32, 32) and bool((a == b).all()) and float(a.min()) >= 0 and float(a.max()) <= 1
    assert not bool((ds[1]["GT"] == a).all())
