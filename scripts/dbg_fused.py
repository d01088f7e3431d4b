# SELFC_TC_DBG=1: CTA 0's per-role barrier-wait cycles of every dense_fused_kernel launch of one 1080p GOP (bf16 mode)
import ctypes as C, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["SELFC_TC_DBG"] = "1"
from selfc_b200.engine import Engine
from selfc_b200 import _lib, synthetic
dev = torch.device("cuda", 0)
net, _ = synthetic.synthetic_net()
eng = Engine(dev, "bf16"); eng.load_state(net.state_dict())
hh, ww = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (1080, 1920)
x = torch.rand(7, 3, hh, ww, device=dev)
L = _lib.lib()
buf = (C.c_longlong * (17 * 4096))()
L.selfc_debug_read.restype = C.c_int
for it in range(2):
    _, _, lrq = eng.down(x, 7, want_out51=False)
    eng.up(lrq, 7, want_hf=False)
    torch.cuda.synchronize()
    n = L.selfc_debug_read(buf, 4096)
names = ["prod_wait_xempty", "prod_total", "mma_wait_tempty", "mma_wait_xfull", "mma_wait_x1", "mma_wait_x2", "mma_wait_x3", "mma_total",
         "steps", "groups", "epi0_wait_tfull", "epi0_total", "epi0_groups", "epi1_wait_tfull", "epi1_total", "epi1_groups"]
seen = {}
for i in range(n):
    tag = buf[17 * i]
    if tag < 8000000 or tag >= 9000000 or tag in seen: continue
    seen[tag] = 1
    v = [buf[17 * i + 1 + j] for j in range(16)]
    print(f"dual={(tag // 100000) % 10} schedule={(tag // 1000) % 10} nx={tag % 1000}: " + ", ".join(f"{k}={x}" for k, x in zip(names, v)))
    if v[8]:
        st = v[8]
        print(f"    per step: mma_total {v[7] / st:.0f} = issue {(v[7] - v[2] - v[3] - v[4] - v[5] - v[6]) / st:.0f} + wait tempty {v[2] / st:.0f} xfull {v[3] / st:.0f} "
              f"x1 {v[4] / st:.0f} x2 {v[5] / st:.0f} x3 {v[6] / st:.0f} | per group: epilogue team0 busy {(v[11] - v[10]) / max(1, v[12]):.0f} wait_tfull {v[10] / max(1, v[12]):.0f}"
              f"  team1 busy {(v[14] - v[13]) / max(1, v[15]):.0f} wait_tfull {v[13] / max(1, v[15]):.0f} | producer wait_xempty {v[0] / st:.0f} of {v[1] / st:.0f}")
