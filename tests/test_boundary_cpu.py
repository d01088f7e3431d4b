"""CPU-side checks of the drop-in boundary: the C-ABI library loads and exports every symbol include/selfc_b200.h
declares (no compute calls without a GPU), the module tree has the reference's state_dict layout, the options loader
behaves like the reference's, and the product path refuses to run without CUDA (no fallback)."""
import os
import re

import pytest
import torch

from oracle import selfc_oracle as so

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "selfc_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(selfc_[a-z0-9_]+)\s*\(", hdr)))


def test_library_exports_every_declared_symbol():
    from selfc_b200 import _lib, build
    build.build()
    L = _lib.lib()
    names = _declared_symbols()
    assert len(names) >= 18
    for n in names:
        assert hasattr(L, n), f"{n} declared in include/selfc_b200.h but not exported"
        assert n in _lib.SIGNATURES, f"{n} has no ctypes signature in selfc_b200/_lib.py"
    assert L.selfc_version() >= 100
    assert L.selfc_last_error() is not None


def test_module_layout_is_the_reference_state_dict():
    from selfc_b200 import networks, options
    from selfc_b200.engine import PARAM_NAMES
    opt = options.dict_to_nonedict(options.parse(os.path.join(ROOT, "selfc_b200", "configs", "selfc_large_synthetic.yml"), is_train=False))
    assert opt["network_G"]["no_such_key"] is None and opt["is_train"] is False
    assert opt["datasets"]["test_1"]["phase"] == "test" and opt["datasets"]["test_1"]["data_type"] == "img"
    net = networks.define_G(opt)
    sd = net.state_dict()
    shapes = so.param_shapes()
    assert list(sd.keys()) == list(shapes.keys()) == list(PARAM_NAMES)
    assert all(tuple(sd[k].shape) == shapes[k] for k in shapes)
    assert sum(v.numel() for v in sd.values()) == 3_365_038
    # strict load of a reference-layout checkpoint, with and without DataParallel's 'module.' prefix stripped by the caller
    net.load_state_dict(so.make_state_dict(0), strict=True)
    assert len(list(net.buffers())) == 0


def test_unsupported_variants_fail_loudly():
    from selfc_b200 import networks, options
    opt = options.dict_to_nonedict(options.parse(os.path.join(ROOT, "selfc_b200", "configs", "selfc_large_synthetic.yml"), is_train=False))
    opt["model"] = "IRN"
    with pytest.raises(NotImplementedError):
        networks.define_G(opt)
    opt["model"] = "SelfC_GMM"
    opt["network_G"]["global_module"] = "nolocal"     # the typo in two reference YAMLs
    with pytest.raises(NotImplementedError):
        networks.define_G(opt)


def test_no_cpu_fallback():
    from selfc_b200 import networks, options
    from selfc_b200.global_var import GlobalVar
    opt = options.dict_to_nonedict(options.parse(os.path.join(ROOT, "selfc_b200", "configs", "selfc_large_synthetic.yml"), is_train=False))
    net = networks.define_G(opt)
    GlobalVar.set_Temporal_LEN(1)
    with pytest.raises(RuntimeError, match="no CPU"):
        net(torch.zeros(1, 3, 8, 8))
    from selfc_b200 import engine
    with pytest.raises(RuntimeError, match="no CPU"):
        engine.fa_forward(torch.zeros(1, 3, 8, 8))


def test_product_never_imports_the_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "selfc_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in src and "from oracle" not in src and "oracle/" not in src, f


@pytest.mark.skipif(not os.path.isdir("/root/reference/codes"), reason="reference mount absent (GPU box)")
def test_same_seed_same_initial_weights_as_reference():
    """torch.manual_seed(s); define_G(opt) draws the same default-init weights as the reference."""
    from oracle import ref_shim
    from selfc_b200 import networks, options
    ref = ref_shim.load_reference()
    torch.manual_seed(3)
    rnet = ref.build_net(7)
    opt = options.dict_to_nonedict(options.parse(ref_shim.VID4_YAML, is_train=False))   # the UNMODIFIED reference YAML
    torch.manual_seed(3)
    mine = networks.define_G(opt)
    a, b = rnet.state_dict(), mine.state_dict()
    assert list(a.keys()) == list(b.keys())
    assert all(torch.equal(a[k], b[k]) for k in a)


def test_packaged_seeded_weights_equal_the_oracle_recipe():
    """bench.py's product arm draws its weights from selfc_b200.synthetic (no oracle import on that path); the recipe is the
    oracle's, tensor for tensor."""
    import torch
    from oracle import selfc_oracle as so
    from selfc_b200.synthetic import seeded_state_dict, synthetic_net
    net, _ = synthetic_net()
    a, b = seeded_state_dict(net, 0), so.make_state_dict(0)
    assert list(a.keys()) == list(b.keys())
    assert all(torch.equal(a[k], b[k]) for k in a)


def test_planar_gmm_layout_helpers_match_the_oracle_sampler():
    """Host logic of selfc_gmm_sample_planar's Python wrapper: the permutation from the reference's parameter order
    (hf*15 + k*3 + j, SelfC_GMM_arch_inv.py:383-388) to the planar quads [j*60 + k*12 + i][M][4] and back from the planar
    latent.  A plain-torch evaluation of the sampler ON the planar layout must reproduce the oracle's draw."""
    from selfc_b200 import engine
    b, t, h, w = 2, 3, 4, 5
    gen = torch.Generator().manual_seed(11)
    params = torch.randn(b * t, 720, h, w, generator=gen)
    eps = so.make_eps(b, t, h, w, 3)
    ref = so.gmm_sample(params, eps, t)
    m = b * t * h * w
    planar = engine.gmm_params_to_planar(params)
    assert planar.shape == (180, m, 4) and planar.is_contiguous()
    # spot check of the index map: quad j*60 + k*12 + i, element e  <-  channel (4i+e)*15 + k*3 + j of pixel m
    for (j, k, i, e, mm) in [(0, 0, 0, 0, 0), (1, 3, 7, 2, 17), (2, 4, 11, 3, m - 1)]:
        n, pix = divmod(mm, h * w)
        assert planar[j * 60 + k * 12 + i, mm, e] == params[n, (4 * i + e) * 15 + k * 3 + j, pix // w, pix % w]
    lg, ls, mu = (planar[j * 60:(j + 1) * 60].reshape(5, 12, m, 4).permute(0, 2, 1, 3).reshape(5, m, 48) for j in range(3))
    pi = torch.softmax(lg, dim=2)                                    # over the 48 HF channels, per component (F3)
    e = eps.permute(2, 0, 3, 4, 5, 1).reshape(5, m, 48)             # [B,48,5,T,h,w] -> [k, m, hf]
    v = (pi * (e * torch.exp(ls.clamp(-7, 7)) + mu)).sum(0)          # [m, 48]
    z = torch.zeros(13, m, 4)
    z[1:] = v.reshape(m, 12, 4).permute(1, 0, 2)
    got = engine.gmm_latent_from_planar(z, b * t, h, w)
    torch.testing.assert_close(got, ref, rtol=1e-5, atol=1e-5)


def test_dense_fused_schedules_keep_every_row_alive_until_its_last_reader():
    """The fused dense-block kernel (csrc/dense_fused.cu) keeps rows of x1..x4 in tensor-memory rings and rows of X in a
    shared-memory ring while four (five) layers walk down a strip, one row per step, `lag` rows apart and in a fixed issue
    `order`.  A ring slot may be overwritten once every reader of its row has been ISSUED before the overwriting row's own MMAs
    (tcgen05.commit covers everything issued earlier), a consumer must be issued after its producer, and everything has to fit 512
    tensor-memory columns.  Simulate the issue order of the COMPILED tables (host-only C-ABI call) and check all three."""
    import ctypes as C
    from selfc_b200 import _lib, build
    build.build()
    L_ = _lib.lib()
    out = (C.c_int * 18)()
    assert L_.selfc_dense_fused_schedule(7, out) != 0
    for sch in range(4):
        assert L_.selfc_dense_fused_schedule(sch, out) == 0
        v = list(out)
        nl, ng, nxr = v[0], v[1], v[2]
        lag, order, ring, cols = v[3:3 + ng], v[8:8 + ng], v[13:17], v[17]
        assert sorted(order) == list(range(ng)) and cols <= 512 and lag[0] == 0
        f5 = ng > nl

        def reads(j, r):          # (source, row): source -1 = X, 0.. = x1..
            if j < nl:            # conv layer j: rows r-1, r, r+1 of X and of every earlier layer
                return [(g, rr) for g in range(-1, j) for rr in (r - 1, r, r + 1) if rr >= 0]
            return [(g, r) for g in range(-1, nl)]      # conv5 taps: row r of everything
        issued = [(s, j, s - lag[j]) for s in range(80) for j in order if s - lag[j] >= 0]
        idx = {(j, r): k for k, (s, j, r) in enumerate(issued)}
        kept = nl if f5 else nl - 1
        for k, (s, j, r) in enumerate(issued):
            for g, rr in reads(j, r):
                if g < 0 or (g, rr) not in idx:
                    continue
                assert g < kept and ring[g] > 0, (sch, "a layer reads rows that are not kept on chip", j, g)
                assert idx[(g, rr)] < k, (sch, "consumer issued before its producer", (j, r), (g, rr))
                over = (g, rr + ring[g])            # the row that re-uses this slot
                assert over not in idx or idx[over] > k, (sch, "ring too small", g, rr, (j, r))
        # X rows alive at once (+1 in flight) must fit the ring
        live = 0
        for s in range(30, 80):
            rows = [rr for j in order for g, rr in reads(j, s - lag[j]) if g == -1]
            live = max(live, max(rows) - min(rows) + 1)
        assert live + 1 <= nxr, (sch, live, nxr)


def test_precision_modes_match_the_header():
    """The three precision modes of the C-ABI (include/selfc_b200.h) and the strings the host side accepts for them."""
    import re
    from selfc_b200 import _lib
    from selfc_b200.engine import parse_mode
    here = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    hdr = open(os.path.join(here, "include", "selfc_b200.h")).read()
    consts = {m.group(1): int(m.group(2)) for m in re.finditer(r"#define (SELFC_MODE_\w+) (\d+)", hdr)}
    assert consts == {"SELFC_MODE_FP32": _lib.MODE_FP32, "SELFC_MODE_BF16": _lib.MODE_BF16, "SELFC_MODE_BF16X3": _lib.MODE_BF16X3}
    assert parse_mode("fp32") == parse_mode(None) == _lib.MODE_FP32
    assert parse_mode("bf16") == _lib.MODE_BF16
    assert parse_mode("bf16x3") == parse_mode("fp32_tc") == _lib.MODE_BF16X3
    with pytest.raises(ValueError):
        parse_mode("tf32")


@pytest.mark.parametrize("B,T,h,w", [(1, 1, 1, 1), (2, 3, 5, 7), (1, 7, 9, 14), (3, 2, 4, 30), (1, 2, 3, 6)])
def test_wgrad_planes_make_every_tap_a_plain_offset(B, T, h, w):
    """The tensor-core weight-gradient kernel (BF16X3 training) contracts over zero-padded pixel planes in which a tap is a plain
    offset of the pixel index.  From the geometry the library reports (no device work): all real pixels get distinct indices inside
    the plane, the row pitch lets TMA boxes start on 16-byte boundaries, and every spatial / temporal neighbour of a real pixel is
    either the true neighbour (inside the frame / the clip) or a padding position -- never a real pixel of another row, frame or clip
    -- including the one-pixel-shifted copies of the gradient plane."""
    import ctypes as C
    import numpy as np
    from selfc_b200 import _lib
    out = (C.c_longlong * 4)()
    assert _lib.lib().selfc_wgrad_geometry(B, T, h, w, out) == 0
    Wp, Fp, P, Pa = [int(v) for v in out]
    assert Wp % 8 == 0 and Wp >= w + 2 and Fp == (h + 2) * Wp and P == (B * (T + 1) + 1) * Fp and Pa % 32 == 0 and Pa >= P
    b, t, y, x = np.meshgrid(np.arange(B), np.arange(T), np.arange(h), np.arange(w), indexing="ij")
    idx = ((b * (T + 1) + t + 1) * (h + 2) + y + 1) * Wp + x + 1
    assert idx.min() >= 0 and idx.max() < P and np.unique(idx).size == idx.size
    real = np.full(Pa + 2 * Fp, -1, dtype=np.int64)            # plane position -> flat pixel id (or -1: padding), with slack either side
    flat = np.arange(idx.size).reshape(idx.shape)
    real[idx + Fp] = flat
    for ky in range(3):
        for kx in range(3):
            got = real[idx + Fp + (ky - 1) * Wp + (kx - 1)]
            yy, xx = y + ky - 1, x + kx - 1
            inside = (yy >= 0) & (yy < h) & (xx >= 0) & (xx < w)
            want = np.where(inside, flat[b, t, np.clip(yy, 0, h - 1), np.clip(xx, 0, w - 1)], -1)
            assert np.array_equal(got, want), (ky, kx)
    for dt in range(3):
        got = real[idx + Fp + (dt - 1) * Fp]
        tt = t + dt - 1
        inside = (tt >= 0) & (tt < T)
        want = np.where(inside, flat[b, np.clip(tt, 0, T - 1), y, x], -1)
        assert np.array_equal(got, want), dt


def test_bf16x3_arithmetic_restated_on_the_host():
    """What the BF16X3 mode computes, restated with torch CPU ops: hi = bf16(v), lo = bf16(v - hi), A.W ~ A_hi.W_hi + A_hi.W_lo + A_lo.W_hi
    with fp32 accumulation.  The split carries 16 mantissa bits (|v - hi - lo| <= 2^-17 |v|) and a K = 1440 dot product (the widest
    (1,3,3) layer: 9 x 160) lands within a few 1e-6 of the fp64 result relative to the operand scale -- three orders of magnitude
    inside the 1e-3 gate, against ~2e-3 for plain bf16 operands."""
    g = torch.Generator().manual_seed(0)
    a = torch.randn(4096, 1440, generator=g)
    w = torch.randn(1440, 32, generator=g) * 0.05

    def split(v):
        hi = v.to(torch.bfloat16).to(torch.float32)
        lo = (v - hi).to(torch.bfloat16).to(torch.float32)
        return hi, lo

    a_hi, a_lo = split(a)
    w_hi, w_lo = split(w)
    assert ((a - a_hi - a_lo).abs() <= a.abs() * 2.0 ** -17 + 1e-30).all()
    ref = a.double() @ w.double()
    x3 = (a_hi @ w_hi + a_hi @ w_lo + a_lo @ w_hi).double()
    bf = (a_hi @ w_hi).double()
    scale = float(ref.abs().max())
    err_x3, err_bf = float((x3 - ref).abs().max()) / scale, float((bf - ref).abs().max()) / scale
    assert err_x3 <= 1e-5 and err_bf >= 50 * err_x3, (err_x3, err_bf)


def test_dependent_launch_kernels_wait_for_their_predecessor():
    """Source lint for the programmatic-dependent-launch discipline (csrc/tc_ptx.cuh `launch_pdl*`, csrc/wgrad_tc.h `launch_chain`): a
    kernel launched with the stream-serialisation attribute may start before the previous kernel of the stream has finished, so its body
    must execute `griddepcontrol.wait` (`pdl_wait()` / `chain_entry()`) -- a kernel without it would read its predecessor's output early
    and no parity test is guaranteed to catch the race."""
    import glob
    import re
    here = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    src = {p: open(p).read() for p in glob.glob(os.path.join(here, "selfc_b200", "csrc", "*.cu"))}
    launched = set()
    for text in src.values():
        for m in re.finditer(r"launch_(?:chain|pdl|pdl_pairs)\(([^;]*?);", text, re.S):
            launched.update(re.findall(r"(\w+_kernel)\b", m.group(1)))
    launched.add("dense_fused_kernel")            # launched through a function pointer chosen by template arguments
    assert {"conv3x3_tc3_kernel", "temporal_tc_kernel", "wgrad_tc_kernel", "wg_planes_grad_kernel", "wg_planes_act_kernel",
            "lrelu_bwd_kernel", "wgrad_unpack_kernel", "cols_to_slab_kernel"} <= launched, launched
    for name in sorted(launched):
        bodies = []
        for text in src.values():
            for m in re.finditer(r"__global__[^;{]*?\b" + name + r"\s*\(", text, re.S):
                end = text.find("\n}\n", m.end())
                bodies.append(text[m.end():end])
        assert bodies, f"no definition of {name} found"
        for body in bodies:
            assert "pdl_wait()" in body or "chain_entry()" in body, f"{name} is launched as a dependent launch but never waits"
