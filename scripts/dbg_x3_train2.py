import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import selfc_oracle as so
from selfc_b200.engine import Engine
dev = torch.device("cuda", 0)
sd = so.make_state_dict(3, gain=1.5)
e32 = Engine(dev, "fp32"); e32.load_state(sd)
ex3 = Engine(dev, "bf16x3"); ex3.load_state(sd)
for (b, t, h, w, blk) in [(2, 3, 9, 12, 4), (2, 3, 11, 14, 4), (1, 1, 8, 8, 4), (1, 7, 16, 16, 4), (2, 3, 9, 12, 0), (1, 3, 9, 12, 4), (2, 1, 9, 12, 4), (1, 1, 9, 12, 4), (1, 1, 30, 30, 4), (1, 1, 31, 31, 4)]:
    gen = torch.Generator().manual_seed(5)
    z = torch.randn(b * t, 51, h, w, generator=gen) * 0.5
    gz = torch.randn(b * t, 51, h, w, generator=gen)
    for rev in (False, True):
        a, ga = e32.invblock_backward(blk, rev, z.to(dev), gz.to(dev), t)
        bb, gb = ex3.invblock_backward(blk, rev, z.to(dev), gz.to(dev), t)
        bad = [k.split(".", 2)[2] for k in ga if (ga[k] - gb[k]).abs().max().item() > 1e-3 * ga[k].abs().max().item() + 1e-5]
        print((b, t, h, w, blk), "rev" if rev else "fwd", "gz err %.1e" % (a - bb).abs().max().item(), "bad:", ",".join(bad) if bad else "-", flush=True)

# ---- the same sensitivity inside the fp32-FMA mode: perturb the block input by 1e-5 (what BF16X3's forward differs by)
print("---- fp32 mode, input perturbed by 1e-5 * randn")
for (b, t, h, w, blk) in [(2, 3, 9, 12, 4), (2, 3, 11, 14, 4), (1, 3, 9, 12, 4), (1, 1, 30, 30, 4), (2, 3, 9, 12, 0)]:
    gen = torch.Generator().manual_seed(5)
    z = torch.randn(b * t, 51, h, w, generator=gen) * 0.5
    gz = torch.randn(b * t, 51, h, w, generator=gen)
    zp = z + 1e-5 * torch.randn(z.shape, generator=gen)
    for rev in (False, True):
        a, ga = e32.invblock_backward(blk, rev, z.to(dev), gz.to(dev), t)
        bb, gb = e32.invblock_backward(blk, rev, zp.to(dev), gz.to(dev), t)
        bad = [k.split(".", 2)[2] for k in ga if (ga[k] - gb[k]).abs().max().item() > 1e-3 * ga[k].abs().max().item() + 1e-5]
        rl2 = max(((ga[k] - gb[k]).double().norm() / ga[k].double().norm()).item() for k in ga)
        print((b, t, h, w, blk), "rev" if rev else "fwd", "gz err %.1e" % (a - bb).abs().max().item(), "max rel L2 %.1e" % rl2, "bad:", ",".join(bad) if bad else "-", flush=True)
