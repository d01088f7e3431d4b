# BF16X3 mode under compute-sanitizer: one small rescale (layer-by-layer (hi, lo) convs, S / Y2 couplings, per-component head) and one
# training step's gradients (tensor-core input / weight gradients, plane builders, accumulate epilogue)
import sys, torch
sys.path.insert(0, "/root/repo")
from oracle import selfc_oracle as so
from selfc_b200.engine import Engine
dev = torch.device("cuda", 0)
eng = Engine(dev, "bf16x3"); eng.load_state(so.make_state_dict(3))
b, t, hh, ww = 1, 3, 76, 132
x = so.make_frames(b, t, hh, ww, 5).to(dev)
lr, rec = eng.rescale(x, t, seed=7, offset=1)
torch.cuda.synchronize()
print("rescale ok", float(rec.abs().mean()))
b, t, hh, ww = 2, 3, 40, 56
x = so.make_frames(b, t, hh, ww, 6)
ref_l = so.gaussian_downsample(x)
grads, losses = eng.train_grads(x.to(dev), ref_l.to(dev), t, seed=3, offset=0)
torch.cuda.synchronize()
print("train ok", float(losses[0]))
