# one small bf16 rescale (fused dense blocks, conv5-taps path, dual G+H launches) for compute-sanitizer
import sys, torch
sys.path.insert(0, "/root/repo")
from oracle import selfc_oracle as so
from selfc_b200.engine import Engine
dev = torch.device("cuda", 0)
eng = Engine(dev, "bf16"); eng.load_state(so.make_state_dict(3))
b, t, hh, ww = 1, 3, 148, 484
x = so.make_frames(b, t, hh, ww, 5).to(dev)
lr, rec = eng.rescale(x, t, seed=7, offset=1)
torch.cuda.synchronize()
print("ok", float(rec.abs().mean()))
