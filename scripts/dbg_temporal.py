import ctypes as C, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["SELFC_TC_DBG"] = "1"
from oracle import selfc_oracle as so
from selfc_b200.engine import Engine
from selfc_b200 import _lib
dev = torch.device("cuda", 0)
eng = Engine(dev, "bf16"); eng.load_state(so.make_state_dict(0))
x = torch.rand(7, 3, 1080, 1920, device=dev)
for _ in range(2):
    _, _, lrq = eng.down(x, 7, want_out51=False)
    eng.up(lrq, 7, want_hf=False)
torch.cuda.synchronize()
L = _lib.lib()
buf = (C.c_longlong * (17 * 4096))()
L.selfc_debug_read.restype = C.c_int
n = L.selfc_debug_read(buf, 4096)
names = ["prod_wait_empty", "prod_total", "mma_wait_full", "mma_wait_tempty", "mma_total", "mma_wait_w", "epi_wait_tfull", "epi_wait_cp", "epi_total", "tiles"]
seen = {}
for i in range(n // 2, n):       # second pass (warm)
    tag = buf[17 * i]
    if tag in seen: continue
    seen[tag] = 1
    vals = [buf[17 * i + 1 + j] for j in range(10)]
    print(f"epi={tag // 1000000} npad={(tag // 1000) % 1000} nks={tag % 1000}: " + ", ".join(f"{k}={v}" for k, v in zip(names, vals)))
