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

seen = {}
for i in range(n):
    tag = buf[17 * i]
    if tag < 8000000 or tag >= 9000000 or tag in seen: continue
    seen[tag] = 1
    vals = [buf[17 * i + 1 + j] for j in range(16)]
    print(tag, "per-warp (warps 2..9) cycles per group: bar+math", vals[:8], " ld..before-bar", vals[8:])
