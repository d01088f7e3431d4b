# bf16-mode error against the oracle as the weight scale grows (stand-in for "trained" weights with larger dynamic range)
import sys, torch
sys.path.insert(0, "/root/repo")
from oracle import selfc_oracle as so
from selfc_b200.engine import Engine
dev = torch.device("cuda", 0)
b, t, hh, ww = 1, 7, 96, 160
x = so.make_frames(b, t, hh, ww, 77)
eps = so.make_eps(b, t, hh // 4, ww // 4, 5)
for gain in (1.0, 1.5, 2.0, 2.5, 3.0):
    sd = so.make_state_dict(0, gain)
    with torch.no_grad():
        z = so.net_down(sd, x, t)
        lr = so.quantize(z[:, :3])
        hr_ref, _ = so.net_up(sd, lr, eps, t)
    res = []
    for mode in ("fp32", "bf16"):
        eng = Engine(dev, mode); eng.load_state(sd)
        _, lr_u8, _ = eng.down(x.to(dev), t, want_out51=False)
        d = (lr_u8.cpu().int() - so.quantize_u8(z[:, :3]).int()).abs()
        hr, _ = eng.up(lr.to(dev), t, eps=eps.to(dev), want_hf=False)
        res.append((mode, d.max().item(), round((d == 0).float().mean().item(), 4), round((d <= 1).float().mean().item(), 5), float((hr.cpu() - hr_ref).abs().max())))
    print("gain", gain, "| latent max", float(z.abs().max()), "HR range", float(hr_ref.min()), float(hr_ref.max()), "|", res)
