# compare the conv5-taps path (default) with SELFC_F5=0 on the fp32 latent after the down pass and after single blocks
import os, subprocess, sys, torch
here = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SNIP = r"""
import sys, torch
sys.path.insert(0, %r)
from oracle import selfc_oracle as so
from selfc_b200.engine import Engine
dev = torch.device("cuda", 0)
eng = Engine(dev, "bf16"); eng.load_state(so.make_state_dict(3))
b, t, hh, ww = [int(v) for v in sys.argv[2:6]]
x = so.make_frames(b, t, hh, ww, 5).to(dev)
out51, lr_u8, lr_q = eng.down(x, t)
torch.save({"out51": out51.cpu(), "lr": lr_u8.cpu()}, sys.argv[1])
""" % here
outs = []
for i, extra in enumerate(({}, {"SELFC_F5": "0"})):
    e = dict(os.environ); e.pop("SELFC_F5", None); e.update(extra)
    out = f"/tmp/f5_{i}.pt"
    r = subprocess.run([sys.executable, "-c", SNIP, out] + sys.argv[1:5], env=e, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-1500:]
    outs.append(torch.load(out))
a, b_ = outs[0]["out51"], outs[1]["out51"]
d = (a - b_).abs()
print("latent LR part: max |diff|", d[:, :3].max().item(), "mean", d[:, :3].mean().item(), " HF part: max", d[:, 3:].max().item(), "mean", d[:, 3:].mean().item())
print("LR magnitude mean", b_[:, :3].abs().mean().item())
dl = (outs[0]["lr"].int() - outs[1]["lr"].int()).abs()
print("codes equal", (dl == 0).float().mean().item())
# where are the big differences? per-frame, per-row
dd = d[:, :3].amax(dim=1)
print("per-frame max", dd.flatten(1).max(dim=1).values.tolist())
print("row max (frame 0)", [round(v, 7) for v in dd[0].max(dim=1).values.tolist()])
print("col max (frame 0)", [round(v, 7) for v in dd[0].max(dim=0).values.tolist()])
