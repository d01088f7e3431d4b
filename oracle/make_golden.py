"""Generate tests/golden/*.npz from the LIVE, UNMODIFIED reference.  Build-container only.

    python -m oracle.make_golden

The reference has no tests or golden vectors; these fixtures are outputs of
the reference's own modules (imported from /root/reference through oracle/ref_shim.py) on
seeded inputs, seeded reference-layout weights (oracle.selfc_oracle.make_state_dict) and an
injected eps.  They pin oracle/selfc_oracle.py, which in turn is what the CUDA path is
checked against on the GPU box (where /root/reference does not exist).
"""
from __future__ import annotations

import hashlib
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))

from oracle import ref_shim, selfc_oracle as so  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")


def sd_digest(sd) -> str:
    h = hashlib.sha256()
    for k, v in sd.items():
        h.update(k.encode())
        h.update(v.numpy().tobytes())
    return h.hexdigest()


def case_net(ref, name, b, t, hh, ww, wseed, xseed, gain=1.0):
    sd = so.make_state_dict(wseed, gain)
    net = ref.build_net(t)
    net.load_state_dict(sd, strict=True)          # the reference's own strict load
    x = so.make_frames(b, t, hh, ww, xseed)
    h, w = hh // 4, ww // 4
    eps = so.make_eps(b, t, h, w, xseed + 7)       # regenerated from the seed by the tests (not stored)
    with torch.no_grad():
        # per-stage capture: SelfCInvNet calls op.forward(...) directly (:456,:487), which bypasses
        # forward hooks, so the instances' forward attributes are wrapped instead.
        stages = {}

        def tap(op, key):
            orig = op.forward

            def wrapped(*a, **k):
                o = orig(*a, **k)
                stages[key + ("_rev" if (len(a) > 1 and a[1]) or k.get("rev") else "")] = o.clone()
                return o
            op.forward = wrapped
        for i, op in enumerate(net.operations):
            tap(op, f"op{i}")
        out, loss_c = net(x)
        from models.modules.Quantization import Quantization   # reference module
        lr = Quantization()(out[:, :3])
        ref.inject_eps(net, eps)
        up_stages = {}

        def grab(key):
            def hook(m, a, o):          # must return None: a returned tensor would REPLACE the output
                up_stages[key] = o.clone()
            return hook
        hooks = [net.stp_net.other_stp_modules.register_forward_hook(grab("feat")),
                 net.stp_net.global_m1.register_forward_hook(grab("ga1")),
                 net.stp_net.local_m1.register_forward_hook(grab("lm1"))]
        hr, hf = net(x=lr, rev=True)
        for hk in hooks:
            hk.remove()
        params = net.stp_net.parameters            # [B,720,T,h,w] (attribute shadowing)
        params = params.transpose(1, 2).reshape(b * t, 720, h, w)
    np.savez_compressed(
        os.path.join(OUT, name + ".npz"),
        meta=np.array([b, t, hh, ww, wseed, xseed], dtype=np.int64), gain=np.float32(gain),
        sd_sha256=np.array(sd_digest(sd)),
        x=x.numpy(), eps_seed=np.int64(xseed + 7), eps_sum=np.float64(eps.double().sum().item()),
        down_fa=stages["op0"].numpy(), down_blk1=stages["op1"].numpy(), down_out=out.numpy(),
        loss_c=np.float32(loss_c.item()),
        lr=lr.numpy(),
        stp_lm1=up_stages["lm1"].numpy(), stp_ga1=up_stages["ga1"].numpy(), stp_feat=up_stages["feat"].numpy(),
        params_sub=params[:, ::16].contiguous().numpy(),
        hf=hf.numpy(), up_after_blk8=stages["op8_rev"].numpy(),
        up_after_blk1=stages["op1_rev"].numpy(), hr=hr.numpy())
    print(name, "down", tuple(out.shape), "hr", tuple(hr.shape),
          "lr range", float(out[:, :3].min()), float(out[:, :3].max()),
          "hr err vs x", float((hr - x).abs().max()))


def case_fa(ref):
    g = torch.Generator().manual_seed(3)
    x = torch.rand(3, 3, 16, 24, generator=g)
    z = torch.randn(3, 51, 4, 6, generator=g)
    fa = ref.arch.FrequencyAnalyzer(3)
    with torch.no_grad():
        np.savez_compressed(os.path.join(OUT, "fa.npz"), x=x.numpy(), fwd=fa(x).numpy(),
                            z=z.numpy(), rev=fa(z, rev=True).numpy())


def case_metrics(ref):
    import data.util as dutil                      # reference modules
    import utils.util as uutil
    from models.Guassian import Guassian_downsample
    g = torch.Generator().manual_seed(5)
    a = torch.rand(4, 3, 40, 56, generator=g)
    bb = (a + 0.05 * torch.randn(4, 3, 40, 56, generator=g)).clamp(0, 1)
    ya, yb = dutil.rgb_to_ycbcr(a), dutil.rgb_to_ycbcr(bb)
    # calculate_psnr / calculate_ssim hard-code .cuda(0) (utils/util.py:206-207,602-603): the
    # arithmetic they wrap is called directly instead - ssim() as-is, psnr by its one-line formula.
    ssims = [float(uutil.ssim(ya[i:i + 1], yb[i:i + 1], data_range=1.0)) for i in range(4)]
    psnrs = [20.0 * torch.log10(1.0 / torch.sqrt(torch.mean((ya[i] - yb[i]) ** 2.0))).item() for i in range(4)]
    lr_ref = Guassian_downsample(a.transpose(0, 1)).transpose(0, 1)
    np.savez_compressed(os.path.join(OUT, "metrics.npz"), a=a.numpy(), b=bb.numpy(), ya=ya.numpy(),
                        ssim=np.array(ssims), psnr=np.array(psnrs), lr_ref=lr_ref.numpy())


def case_u8(ref):
    """8-bit ingest/egress: the reference's own read_img1 on a PNG written by cv2, the dataset's two conversion lines
    (LQGTVID_dataset.py:150-154, applied here verbatim because the class needs a directory tree), and tensor2img."""
    import tempfile
    import cv2
    import data.util as dutil                      # reference modules
    import utils.util as uutil
    rng = np.random.default_rng(9)
    img = rng.integers(0, 256, size=(2, 12, 16, 3), dtype=np.uint8)
    img[0, 0, :, 0] = np.arange(16) * 17            # every 17th code incl. 0 and 255
    xs = []
    with tempfile.TemporaryDirectory() as td:
        for i in range(2):
            path = os.path.join(td, f"{i}.png")
            cv2.imwrite(path, img[i])
            im = dutil.read_img1(None, path)
            if im.shape[2] == 3:
                im = im[:, :, [2, 1, 0]]
            xs.append(torch.from_numpy(np.ascontiguousarray(np.transpose(im, (2, 0, 1)))).float())
    from_ref = torch.stack(xs)
    g = torch.Generator().manual_seed(10)
    y = torch.rand(3, 3, 8, 12, generator=g) * 1.3 - 0.15          # values outside [0,1] on both sides
    ties = (torch.arange(0, 96, dtype=torch.float32) * 2.5 + 0.5) / 255.0   # k+0.5 ties, even and odd k
    y[0, 0].view(-1)[:96] = ties
    y[1, 1].view(-1)[:96] = torch.arange(0, 96, dtype=torch.float32) / 255.0 * 2.6
    to_ref = np.stack([uutil.tensor2img(y[i]) for i in range(3)])
    np.savez_compressed(os.path.join(OUT, "u8.npz"), img=img, from_ref=from_ref.numpy(), y=y.numpy(), to_ref=to_ref)


def case_2x(ref):
    """f3: the 2x operators of the sibling configurations, from the reference's own modules."""
    import models.modules.SelfC_arch_inv as haar_mod              # reference modules
    import models.modules.SelfC_Codec_arch_inv as codec_mod
    g = torch.Generator().manual_seed(17)
    x = torch.rand(2, 3, 12, 20, generator=g)
    z15 = torch.randn(2, 15, 6, 10, generator=g)
    xh = torch.rand(2, 5, 8, 12, generator=g)
    zh = torch.randn(2, 12, 4, 6, generator=g)
    fa = codec_mod.FrequencyAnalyzer(3)                            # k defaults to 2
    h3, h5 = haar_mod.HaarDownsampling(3), haar_mod.HaarDownsampling(5)
    with torch.no_grad():
        np.savez_compressed(os.path.join(OUT, "ops2x.npz"), x=x.numpy(), fa2_fwd=fa(x).numpy(), z15=z15.numpy(),
                            fa2_rev=fa(z15, rev=True).numpy(), xh=xh.numpy(), haar5_fwd=h5(xh).numpy(),
                            haar3_fwd=h3(x).numpy(), zh=zh.numpy(), haar3_rev=h3(zh, rev=True).numpy(),
                            haar_state_keys=np.array(sorted(h3.state_dict().keys())))


def case_sampler(ref):
    """f4: index streams of the reference's DistIterSampler (data/data_sampler.py) for a few (size, world, ratio, epoch)."""
    from data.data_sampler import DistIterSampler              # reference module
    out = {}
    for size, world, ratio in ((37, 3, 4), (64, 2, 200), (5, 4, 1)):
        ds = list(range(size))
        for rank in range(world):
            smp = DistIterSampler(ds, world, rank, ratio)
            for epoch in (0, 3):
                smp.set_epoch(epoch)
                out[f"s{size}_w{world}_r{ratio}_k{rank}_e{epoch}"] = np.array(list(iter(smp)), dtype=np.int32)
                out[f"s{size}_w{world}_r{ratio}_k{rank}_len"] = np.int64(len(smp))
    np.savez_compressed(os.path.join(OUT, "sampler.npz"), torch_version=np.array(torch.__version__), **out)


def case_train(ref, name, b, t, hh, ww, wseed, xseed):
    """One training step's losses and gradients from the reference's own modules (SelfC_model.py:148-170 restated with
    netG, Quantization, ReconstructionLoss and Guassian_downsample imported from the reference; SelfCModel itself needs
    a CUDA device).  Stored: the loss terms, every parameter's gradient norm, a few whole gradients."""
    from models.modules.Quantization import Quantization
    from models.modules.loss import ReconstructionLoss
    from models.Guassian import Guassian_downsample
    sd = so.make_state_dict(wseed)
    net = ref.build_net(t)
    net.load_state_dict(sd, strict=True)
    net.train()
    x = so.make_frames(b, t, hh, ww, xseed)
    eps = so.make_eps(b, t, hh // 4, ww // 4, xseed + 7)
    ref_l = Guassian_downsample(x.transpose(0, 1)).transpose(0, 1)          # distortion: sr_bd (SelfC_model.py:127-128)
    out, loss_c = net(x=x, rev=False)
    lr_pre = out[:, :3]
    l_forw = 1.0 * ReconstructionLoss(losstype="l2")(lr_pre, ref_l.detach())
    lr_q = Quantization()(lr_pre)
    ref.inject_eps(net, eps)
    hr, _ = net(x=lr_q, rev=True)
    l_back = 1.0 * ReconstructionLoss(losstype="l1")(x, hr[:, :3])
    loss = (l_forw + l_back + loss_c.mean() * 0) * 144 * 144 * 3
    loss.backward()
    names = [k for k, _ in net.named_parameters()]
    norms = np.array([float(p.grad.double().norm()) for _, p in net.named_parameters()])
    keep = ["operations.1.F.conv1.weight", "operations.8.G.conv5.bias", "operations.4.H.conv3.weight",
            "stp_net.global_m1.fc.weight", "stp_net.global_m2.proj2.weight", "stp_net.tail_gmm.5.bias",
            "stp_net.local_m1.conv5.weight"]
    grads = {"grad__" + k.replace(".", "__"): dict(net.named_parameters())[k].grad.numpy() for k in keep}
    np.savez_compressed(os.path.join(OUT, name + ".npz"),
                        meta=np.array([b, t, hh, ww, wseed, xseed], dtype=np.int64), eps_seed=np.int64(xseed + 7),
                        names=np.array(names), grad_norms=norms, loss=np.float64(loss.item()),
                        l_forw=np.float64(l_forw.item()), l_back=np.float64(l_back.item()), ref_l=ref_l.numpy(), **grads)
    print(name, "loss", loss.item(), "l_forw", l_forw.item(), "l_back", l_back.item(), "global grad norm", float(np.sqrt((norms ** 2).sum())))


def main():
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(os.cpu_count())
    ref = ref_shim.load_reference()
    case_fa(ref)
    case_metrics(ref)
    case_u8(ref)
    case_2x(ref)
    case_sampler(ref)
    case_net(ref, "net_t3", b=2, t=3, hh=32, ww=48, wseed=0, xseed=11)
    case_net(ref, "net_t7", b=1, t=7, hh=32, ww=40, wseed=1, xseed=12)
    # partial 8x16 output tiles, non-integral 32x32 pooling windows (h=10, w=18), larger weights
    case_net(ref, "net_t2_gain", b=2, t=2, hh=40, ww=72, wseed=2, xseed=13, gain=1.5)
    case_train(ref, "train_t3", b=2, t=3, hh=32, ww=48, wseed=4, xseed=21)


if __name__ == "__main__":
    main()
