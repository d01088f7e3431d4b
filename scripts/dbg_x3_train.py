"""BF16X3 training pieces against the fp32-FMA mode (itself autograd-checked): per-channel errors of an InvBlockExp backward."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import selfc_oracle as so
from selfc_b200.engine import Engine
dev = torch.device("cuda", 0)
sd = so.make_state_dict(3, gain=1.5)
b, t, h, w, blk = 2, 3, 9, 12, 4
gen = torch.Generator().manual_seed(5)
z = torch.randn(b * t, 51, h, w, generator=gen) * 0.5
gz = torch.randn(b * t, 51, h, w, generator=gen)
res = {}
for mode in ("fp32", "bf16x3"):
    eng = Engine(dev, mode); eng.load_state(sd)
    for rev in (False, True):
        gzin, grads = eng.invblock_backward(blk, rev, z.to(dev), gz.to(dev), t)
        res[(mode, rev)] = (gzin.cpu(), {k: v.cpu() for k, v in grads.items()})
for rev in (False, True):
    a, ga = res[("fp32", rev)]; bb, gb = res[("bf16x3", rev)]
    d = (a - bb).abs().amax(dim=(0, 2, 3))
    print("rev", rev, "gzin per-channel max err:", [f"{v:.1e}" for v in d.tolist()])
    for k in ga:
        e = (ga[k] - gb[k]).abs().max().item(); m = ga[k].abs().max().item()
        print(f"   {k}: err {e:.2e} of {m:.2e}" + ("   <<<" if e > 1e-3 * m + 1e-5 else ""))

# ---- G alone at the same shape / data: forward and backward through the component entry points
print("---- G alone")
prefix = "operations.5"
x1, x2 = z[:, :3], z[:, 3:]
with torch.no_grad():
    y1 = x1 + so.d2dt(sd, prefix + ".F", x2, t)
    gref = so.d2dt(sd, prefix + ".G", y1, t)
gy2 = gz[:, 3:].contiguous()
out = {}
for mode in ("fp32", "bf16x3"):
    eng = Engine(dev, mode); eng.load_state(sd)
    f = eng.d2dt(prefix + ".G", y1.to(dev), t).cpu()
    print(mode, "G forward vs oracle:", (f - gref).abs().max().item())
    gx, grads = eng.d2dt_backward(prefix + ".G", y1.to(dev), gy2.to(dev), t)
    out[mode] = (gx.cpu(), {k: v.cpu() for k, v in grads.items()})
a, ga = out["fp32"]; bb, gb = out["bf16x3"]
print("gx err", (a - bb).abs().max().item())
for k in ga:
    e = (ga[k] - gb[k]).abs().max().item(); m = ga[k].abs().max().item()
    print(f"   {k}: err {e:.2e} of {m:.2e}" + ("   <<<" if e > 1e-3 * m + 1e-5 else ""))
