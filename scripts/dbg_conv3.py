# SELFC_TC_DBG=1 + a build with -DSELFC_TC_TIMING: CTA 0's barrier-wait cycles of every conv3x3 launch of one 1080p GOP
import ctypes as C, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["SELFC_TC_DBG"] = "1"
from selfc_b200.synthetic import seeded_state_dict
from selfc_b200.engine import Engine
from selfc_b200 import _lib, synthetic
dev = torch.device("cuda", 0)
net, _ = synthetic.synthetic_net()
eng = Engine(dev, "bf16"); eng.load_state(net.state_dict())
x = torch.rand(7, 3, 1080, 1920, device=dev)
L = _lib.lib()
buf = (C.c_longlong * (17 * 4096))()
L.selfc_debug_read.restype = C.c_int
for it in range(2):
    _, _, lrq = eng.down(x, 7, want_out51=False)
    eng.up(lrq, 7, want_hf=False)
    torch.cuda.synchronize()
    n = L.selfc_debug_read(buf, 4096)
names = ["prod_wait_empty", "prod_total", "mma_wait_full", "mma_wait_tempty", "mma_total", "-", "epi_wait_tfull", "-", "epi_total", "tiles"]
seen = {}
for i in range(n):
    tag = buf[17 * i]
    if tag < 9000000 or tag in seen: continue
    seen[tag] = 1
    vals = [buf[17 * i + 1 + j] for j in range(16)]
    print(f"   pdl_wait max={vals[10]} avg={vals[11] / 148:.0f}  epi_busy max={vals[12]} avg={vals[13] / 148:.0f}  epi_total max={vals[14]} avg={vals[15] / 148:.0f}")
    print(f"dual={(tag // 100000) % 10} nks={tag % 1000}: " + ", ".join(f"{k}={v}" for k, v in zip(names, vals) if k != "-"))
