"""Kernel-time breakdown of one training step with the torch profiler (CUPTI sees every kernel of the process, the library's
included): cheap compared with an ncu launch list.   gpurun -- 'python scripts/prof_train.py [mode] > gpurun_out/prof_train.txt'"""
import collections
import os
import sys

import torch
from torch.profiler import ProfilerActivity, profile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from selfc_b200 import engine as _eng  # noqa: E402
from selfc_b200.global_var import GlobalVar  # noqa: E402
from selfc_b200.synthetic import seeded_state_dict, synthetic_net  # noqa: E402
from selfc_b200.train import Trainer  # noqa: E402

mode = sys.argv[1] if len(sys.argv) > 1 else "bf16x3"
nb = int(sys.argv[2]) if len(sys.argv) > 2 else 1
dev = torch.device("cuda", 0)
t, hh, ww = 7, 256, 448
net, _ = synthetic_net(train=True)
net.load_state_dict(seeded_state_dict(net, 0), strict=True)
net = net.to(dev)
net.set_precision(mode)
GlobalVar.set_Temporal_LEN(t)
tr = Trainer(net, dev, lr=1e-4, weight_decay=1e-14, max_norm=10.0)
x = bench.make_group(nb * t, hh, ww, 4321, dev)
ref_l = _eng.gaussian_downsample(x)
for i in range(2):
    tr.step(x, ref_l, t, seed=42, offset=i)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    tr.step(x, ref_l, t, seed=42, offset=2)
    torch.cuda.synchronize()
agg = collections.OrderedDict()
for e in prof.events():
    if e.device_type is not None and "cuda" in str(e.device_type).lower():
        a = agg.setdefault(e.name.split("(")[0].replace("void ", "").replace("selfc::", "")[:70], [0, 0.0])
        a[0] += 1
        a[1] += e.device_time / 1e3 if hasattr(e, "device_time") else e.cuda_time / 1e3
tot = sum(a[1] for a in agg.values())
print(f"mode {mode}, {nb} septuplet(s): {sum(a[0] for a in agg.values())} kernels, {tot:.2f} ms of kernel time")
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:70]:
    print(f"{k:72s} {a[0]:5d} {a[1]:9.3f} ms {a[1] / tot:.3f}")

# timeline: busy time (union of kernel intervals) against the span of the step, and which kernels the idle gaps follow
ev = []
for e in prof.events():
    if e.device_type is not None and "cuda" in str(e.device_type).lower():
        dur = e.device_time if hasattr(e, "device_time") else e.cuda_time
        ev.append((e.time_range.start, e.time_range.start + dur, e.name.split("(")[0].replace("void ", "").replace("selfc::", "")[:50]))
ev.sort()
span = ev[-1][1] - ev[0][0]
busy, cur_end, gaps = 0.0, ev[0][0], collections.OrderedDict()
for i, (a, b, name) in enumerate(ev):
    if a > cur_end:
        prev = ev[i - 1][2] if i else "-"
        g = gaps.setdefault(prev + " -> " + name, [0, 0.0])
        g[0] += 1
        g[1] += a - cur_end
        busy += b - a
    else:
        busy += max(0.0, b - cur_end)
    cur_end = max(cur_end, b)
print(f"span {span / 1e3:.2f} ms, busy {busy / 1e3:.2f} ms, idle {(span - busy) / 1e3:.2f} ms")
for k, g in sorted(gaps.items(), key=lambda kv: -kv[1][1])[:25]:
    print(f"{k:104s} {g[0]:5d} {g[1] / 1e3:8.3f} ms")
