# device noise stream vs its numpy restatement: max / quantiles of |diff|, and the sampler's time in one 1080p GOP
import sys, numpy as np, torch
sys.path.insert(0, "/root/repo")
from oracle import selfc_oracle as so
from selfc_b200 import engine
from selfc_b200.engine import Engine
dev = torch.device("cuda", 0)
a = engine.export_eps(2, 7, 16, 24, seed=42, offset=0, device=dev).cpu().numpy()
ref = so.philox_eps(2, 7, 16, 24, seed=42, offset=0)
d = np.abs(a - ref)
print("eps vs numpy: max |diff| %.3e, 99.9%% %.3e, 99%% %.3e, mean %.3e; fraction > 2e-6: %.5f" % (d.max(), np.quantile(d, 0.999), np.quantile(d, 0.99), d.mean(), (d > 2e-6).mean()))
eng = Engine(dev, "bf16"); eng.load_state(so.make_state_dict(0))
x = torch.rand(7, 3, 1080, 1920, device=dev)
eng.prof_enable(True)
for _ in range(3):
    eng.rescale(x, 7, seed=1, offset=0)
torch.cuda.synchronize()
p = eng.prof_read()
print("sampler ms per GOP", p["sampler"]["ms"] / 3, " all classes", {k: round(v["ms"] / 3, 3) for k, v in p.items()})
