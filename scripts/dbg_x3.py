"""BF16X3 mode (the numerics-gate mode on tensor cores), stage by stage against the CPU oracle: prints the max |diff| of every
component entry point and of the whole network instead of asserting, so one GPU call localises a fault.
    gpurun -- 'python scripts/dbg_x3.py > gpurun_out/dbg_x3.txt 2>&1'"""
import os
import sys
import time
import traceback

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import selfc_oracle as so          # noqa: E402  (checker only)
from selfc_b200.engine import Engine           # noqa: E402

dev = torch.device("cuda", 0)
MODE = os.environ.get("X3_MODE", "bf16x3")


def stage(name, fn):
    try:
        t0 = time.time()
        fn()
        torch.cuda.synchronize()
        print(f"[ok] {name} ({time.time() - t0:.1f} s)", flush=True)
    except Exception:
        print(f"[FAIL] {name}\n{traceback.format_exc()}", flush=True)


def conv3x3():
    sd = so.make_state_dict(9)
    eng = Engine(dev, MODE)
    eng.load_state(sd)
    for prefix, cin, k in [("operations.5.G", 3, 0), ("operations.2.F", 48, 0), ("operations.2.F", 48, 3), ("stp_net.local_m2", 64, 3)]:
        for (b, t, h, w) in [(1, 2, 13, 21), (2, 3, 40, 70)]:
            x = torch.randn(b * t, cin + 32 * k, h, w, generator=torch.Generator().manual_seed(cin + k)) * 0.7
            wgt, bias = sd[f"{prefix}.conv{k + 1}.weight"], sd[f"{prefix}.conv{k + 1}.bias"]
            ref = F.leaky_relu(F.conv2d(x.double(), wgt[:, :, 0].double(), bias.double(), padding=1), 0.2).float()
            got = eng.conv3x3(prefix, k, x.to(dev), t).cpu()
            print(f"  conv3x3 {prefix} k={k} {b}x{t}x{h}x{w}: max |diff| {(got - ref).abs().max().item():.3e} (ref max {ref.abs().max().item():.2f})", flush=True)


def d2dt():
    sd = so.make_state_dict(5)
    eng = Engine(dev, MODE)
    eng.load_state(sd)
    for prefix, cin in [("operations.1.F", 48), ("operations.3.G", 3), ("stp_net.local_m2", 64), ("stp_net.local_m1", 3)]:
        for (t, h, w) in [(1, 9, 14), (3, 13, 21), (7, 24, 40)]:
            x = torch.randn(2 * t, cin, h, w, generator=torch.Generator().manual_seed(7)) * 0.5
            with torch.no_grad():
                ref = so.d2dt(sd, prefix, x, t)
            got = eng.d2dt(prefix, x.to(dev), t).cpu()
            print(f"  d2dt {prefix} t={t} {h}x{w}: max |diff| {(got - ref).abs().max().item():.3e} (ref max {ref.abs().max().item():.2f})", flush=True)


def global_agg():
    sd = so.make_state_dict(6, gain=2.0)
    eng = Engine(dev, MODE)
    eng.load_state(sd)
    for (h, w, t) in [(10, 18, 2), (45, 67, 7), (40, 52, 12)]:
        x = torch.randn(2 * t, 64, h, w, generator=torch.Generator().manual_seed(h))
        with torch.no_grad():
            ref = so.global_agg(sd, "stp_net.global_m2", x, t)
        got, wmat = eng.global_agg("stp_net.global_m2", x.to(dev), t)
        print(f"  global_agg {h}x{w} t={t}: max |diff| {(got.cpu() - ref).abs().max().item():.3e}", flush=True)


def network():
    sd = so.make_state_dict(0)
    eng = Engine(dev, MODE)
    eng.load_state(sd)
    for (b, t, hh, ww) in [(1, 1, 4, 4), (1, 3, 36, 44), (2, 7, 96, 160), (3, 5, 8, 12), (1, 2, 484, 16), (1, 7, 576, 704)]:
        x = so.make_frames(b, t, hh, ww, 77)
        eps = so.make_eps(b, t, hh // 4, ww // 4, 5)
        torch.set_num_threads(os.cpu_count() or 1)
        with torch.no_grad():
            z = so.net_down(sd, x, t)
            lr = so.quantize(z[:, :3])
            hr_ref, hf_ref = so.net_up(sd, lr, eps, t)
        out51, lr_u8, _ = eng.down(x.to(dev), t)
        d = (lr_u8.cpu().int() - so.quantize_u8(z[:, :3]).int()).abs()
        hr, hf = eng.up(lr.to(dev), t, eps=eps.to(dev))
        print(f"  net {b}x{t}x{hh}x{ww}: latent {(out51.cpu() - z).abs().max().item():.3e}, LR exact {(d == 0).float().mean().item():.6f} max {d.max().item()}, "
              f"hf {(hf.cpu() - hf_ref).abs().max().item():.3e}, HR {(hr.cpu() - hr_ref).abs().max().item():.3e}", flush=True)


def speed():
    sd = so.make_state_dict(0)
    eng = Engine(dev, MODE)
    eng.load_state(sd)
    b, t, hh, ww = 1, 7, 1080, 1920
    x = so.make_frames(b, t, hh, ww, 3).to(dev)
    for it in range(3):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out51, lr_u8, lr_q = eng.down(x, t)
        hr, hf = eng.up(lr_q, t, seed=1, offset=0)
        e1.record()
        torch.cuda.synchronize()
        print(f"  1080p GOP down+up: {e0.elapsed_time(e1):.2f} ms -> {7000.0 / e0.elapsed_time(e1):.1f} frames/s", flush=True)


if __name__ == "__main__":
    which = sys.argv[1:] or ["conv3x3", "d2dt", "global_agg", "network", "speed"]
    for name in which:
        stage(name, globals()[name])
