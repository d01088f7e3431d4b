"""GPU parity tests: the sm_100a kernels (through the C-ABI) against the CPU oracle on the same seeded inputs,
against the committed golden fixtures, and through size-independent properties at the benchmark's full size.

Tolerances (BASELINE.json north_star): LR uint8 within +-1 LSB on >= 99.99 % of pixels; HR max-abs <= 1e-3 in fp32
mode, <= 2e-2 in bf16 mode.  Integer / index work (FrequencyAnalyzer, quantisation) is bit-exact.
"""
import os

import numpy as np
import pytest
import torch

from oracle import selfc_oracle as so

pytestmark = pytest.mark.gpu

HR_TOL_FP32 = 1e-3
HR_TOL_BF16 = 2e-2
# bf16 mode, LR codes: the north star's gate is "within 1 LSB on >= 99.99 %".  The eight F blocks' bf16 rounding (~5e-4 on a
# 3.9e-3 code step) moves ~10 % of the codes across a rounding boundary (measured 89.4 % exact); what must NOT happen is a
# systematic offset, so the exact-match fraction and the mean signed difference are bounded too (and printed).
LR_EXACT_BF16 = 0.80
LR_BIAS_BF16 = 0.05


@pytest.fixture(scope="module")
def dev():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    return torch.device("cuda", 0)


def _engine(dev, sd, mode="fp32"):
    from selfc_b200.engine import Engine
    eng = Engine(dev, mode)
    eng.load_state(sd)
    return eng


def _t(a):
    return torch.from_numpy(np.asarray(a))


# ------------------------------------------------------------------------------------------------ a1 / a10 / a4
def test_fa_golden_bit_exact(dev, golden_dir):
    from selfc_b200 import engine
    g = np.load(os.path.join(golden_dir, "fa.npz"))
    assert torch.equal(engine.fa_forward(_t(g["x"]).to(dev)).cpu(), _t(g["fwd"]))
    assert torch.equal(engine.fa_reverse(_t(g["z"]).to(dev)).cpu(), _t(g["rev"]))


@pytest.mark.parametrize("shape", [(1, 4, 4), (2, 8, 12), (3, 36, 52), (7, 144, 176)])
def test_fa_vs_oracle_bit_exact(dev, shape):
    from selfc_b200 import engine
    n, hh, ww = shape
    gen = torch.Generator().manual_seed(hh * 131 + ww)
    x = torch.rand(n, 3, hh, ww, generator=gen)
    z = torch.randn(n, 51, hh // 4, ww // 4, generator=gen)
    assert torch.equal(engine.fa_forward(x.to(dev)).cpu(), so.fa_forward(x))
    assert torch.equal(engine.fa_reverse(z.to(dev)).cpu(), so.fa_reverse(z))


def test_quantize_bit_exact(dev):
    from selfc_b200 import engine
    k = torch.arange(0, 256, dtype=torch.float32)
    x = torch.cat([k / 255.0, (k + 0.5) / 255.0, (k + 0.49999) / 255.0, torch.tensor([-1.0, -0.0, 1.0, 1.5, 2e-3]),
                   torch.rand(10007, generator=torch.Generator().manual_seed(1)) * 1.2 - 0.1])
    q8, qf = engine.quantize(x.to(dev))
    assert torch.equal(q8.cpu(), so.quantize_u8(x))
    assert torch.equal(qf.cpu(), so.quantize(x))


# ------------------------------------------------------------------------------------------------ a3 / a6 / a7
@pytest.mark.parametrize("prefix,cin", [("operations.1.F", 48), ("operations.3.G", 3), ("operations.8.H", 3),
                                        ("stp_net.local_m1", 3), ("stp_net.local_m2", 64),
                                        ("stp_net.other_stp_modules.4", 64)])
def test_dense_block_vs_oracle(dev, prefix, cin):
    sd = so.make_state_dict(5)
    eng = _engine(dev, sd)
    b, t, h, w = 2, 3, 13, 21     # ragged: not a multiple of any tile
    x = torch.randn(b * t, cin, h, w, generator=torch.Generator().manual_seed(7)) * 0.5
    with torch.no_grad():
        ref = so.d2dt(sd, prefix, x, t)
    got = eng.d2dt(prefix, x.to(dev), t).cpu()
    torch.testing.assert_close(got, ref, rtol=0, atol=2e-5)


@pytest.mark.parametrize("h,w,t", [(10, 18, 2), (32, 32, 3), (45, 67, 7), (40, 52, 12), (64, 96, 5)])
def test_global_agg_vs_oracle(dev, h, w, t):
    sd = so.make_state_dict(6, gain=2.0)
    eng = _engine(dev, sd)
    b = 2
    x = torch.randn(b * t, 64, h, w, generator=torch.Generator().manual_seed(h))
    with torch.no_grad():
        ref = so.global_agg(sd, "stp_net.global_m2", x, t)
        wref = so.global_agg_weights(sd, "stp_net.global_m2", x, t)
    got, wmat = eng.global_agg("stp_net.global_m2", x.to(dev), t)
    torch.testing.assert_close(wmat.cpu(), wref, rtol=0, atol=1e-6)
    torch.testing.assert_close(got.cpu(), ref, rtol=0, atol=2e-5)


def test_gmm_sample_vs_oracle(dev):
    from selfc_b200 import engine
    b, t, h, w = 2, 3, 5, 9
    gen = torch.Generator().manual_seed(3)
    params = torch.randn(b * t, 720, h, w, generator=gen) * 3.0      # exercises the +-7 clamp
    eps = so.make_eps(b, t, h, w, 17)
    ref = so.gmm_sample(params, eps, t)
    got = engine.gmm_sample(params.to(dev), t, eps=eps.to(dev)).cpu()
    if not torch.allclose(got, ref, rtol=1e-5, atol=1e-4):       # diagnose against a float64 evaluation: which side is off?
        ref64 = so.gmm_sample(params.double(), eps.double(), t)
        got_b = engine.gmm_sample(params.to(dev), t, eps=eps.to(dev)).cpu()
        print("gmm_sample mismatch: |gpu-f64| %.3e  |cpu32-f64| %.3e  |gpu-gpu_again| %.3e  cpu32 deterministic %s" % (
            (got.double() - ref64).abs().max().item(), (ref.double() - ref64).abs().max().item(),
            (got - got_b).abs().max().item(), torch.equal(ref, so.gmm_sample(params, eps, t))))
    torch.testing.assert_close(got, ref, rtol=1e-5, atol=1e-4)
    # counter-based noise: export the stream, feed it to the oracle
    eps2 = engine.export_eps(b, t, h, w, seed=42, offset=3, device=dev)
    got2 = engine.gmm_sample(params.to(dev), t, seed=42, offset=3).cpu()
    torch.testing.assert_close(got2, so.gmm_sample(params, eps2.cpu(), t), rtol=1e-5, atol=1e-4)


@pytest.mark.parametrize("form", [0, 1, -1])
def test_gmm_sample_planar_vs_oracle(dev, form):
    """The bf16 mode's sampler (planar parameter quads written by the tcgen05 head) through its own entry point, both kernel
    forms (thread-per-pixel, warp-split) and the default: injected eps and the counter-based stream; ragged pixel counts
    (M % 32 != 0 and M % 128 != 0) exercise the shadow lanes of the warp-split form."""
    from selfc_b200 import engine
    for (b, t, h, w) in [(2, 3, 5, 9), (1, 7, 6, 11), (1, 2, 8, 8)]:
        gen = torch.Generator().manual_seed(5 + h)
        params = torch.randn(b * t, 720, h, w, generator=gen) * 3.0      # exercises the +-7 clamp
        eps = so.make_eps(b, t, h, w, 23)
        ref = so.gmm_sample(params, eps, t)
        got = engine.gmm_sample_planar(params.to(dev), t, eps=eps.to(dev), form=form).cpu()
        torch.testing.assert_close(got, ref, rtol=1e-5, atol=1e-4)
        eps2 = engine.export_eps(b, t, h, w, seed=42, offset=3, device=dev)
        got2 = engine.gmm_sample_planar(params.to(dev), t, seed=42, offset=3, form=form).cpu()
        torch.testing.assert_close(got2, so.gmm_sample(params, eps2.cpu(), t), rtol=1e-5, atol=1e-4)
        # the NCHW kernel of the fp32 mode draws the same numbers
        torch.testing.assert_close(got2, engine.gmm_sample(params.to(dev), t, seed=42, offset=3).cpu(), rtol=1e-5, atol=1e-5)


def test_philox_stream(dev):
    from selfc_b200 import engine
    a = engine.export_eps(1, 7, 16, 24, seed=42, offset=0, device=dev).cpu()
    b = engine.export_eps(1, 7, 16, 24, seed=42, offset=0, device=dev).cpu()
    c = engine.export_eps(1, 7, 16, 24, seed=42, offset=1, device=dev).cpu()
    assert torch.equal(a, b) and not torch.equal(a, c)
    assert abs(a.mean().item()) < 5e-3 and abs(a.std().item() - 1.0) < 5e-3
    # keyed on the linear index: a bigger batch's first clip draws the same numbers (GPU-count invariance)
    big = engine.export_eps(2, 7, 16, 24, seed=42, offset=0, device=dev).cpu()
    assert torch.equal(big[0], a[0])
    # the numpy restatement of the stream (one Philox call -> the four hf of a quad), every element
    ref = so.philox_eps(1, 7, 16, 24, seed=42, offset=0)
    np.testing.assert_allclose(a.numpy(), ref, rtol=0, atol=2e-6)
    ref1 = so.philox_eps(2, 7, 16, 24, seed=42, offset=0)
    np.testing.assert_allclose(big.numpy(), ref1, rtol=0, atol=2e-6)


# ------------------------------------------------------------------------------------------------ a2 / a9 / a11
@pytest.mark.parametrize("mode", ["fp32", "bf16x3"])
@pytest.mark.parametrize("name", ["net_t3", "net_t7", "net_t2_gain"])
def test_network_vs_golden_and_oracle(dev, golden_dir, name, mode):
    """Both numerics-gate modes (fp32-FMA kernels; (hi, lo) bf16 operands on the tensor cores) against the outputs of the LIVE REFERENCE
    (tests/golden/net_*.npz, oracle/make_golden.py): latent, LR codes, HF sample, HR."""
    g = np.load(os.path.join(golden_dir, name + ".npz"))
    b, t, hh, ww, wseed, xseed = [int(v) for v in g["meta"]]
    sd = so.make_state_dict(wseed, float(g["gain"]))
    eng = _engine(dev, sd, mode)
    x = _t(g["x"])
    out51, lr_u8, lr_q = eng.down(x.to(dev), t)
    torch.testing.assert_close(out51.cpu(), _t(g["down_out"]), rtol=0, atol=1e-4 if mode == "fp32" else 4e-4)
    ref_u8 = so.quantize_u8(_t(g["down_out"])[:, :3])
    diff = (lr_u8.cpu().int() - ref_u8.int()).abs()
    exact = (diff == 0).float().mean().item()
    assert diff.max().item() <= 1 and (exact >= 0.9999 or (mode == "bf16x3" and round((1.0 - exact) * diff.numel()) <= 3))
    assert torch.equal(lr_q.cpu(), lr_u8.cpu().float() / 255.0)
    # up, from the REFERENCE's quantised LR so both sides see identical inputs
    eps = so.make_eps(b, t, hh // 4, ww // 4, int(g["eps_seed"]))
    hr, hf = eng.up(_t(g["lr"]).to(dev), t, eps=eps.to(dev))
    torch.testing.assert_close(hf.cpu(), _t(g["hf"]), rtol=0, atol=5e-4)
    torch.testing.assert_close(hr.cpu(), _t(g["hr"]), rtol=0, atol=HR_TOL_FP32)


def test_vid4_shape_clip_vs_oracle(dev):
    """configs[0]/[1]: one 7-frame 576x704 clip, fp32 mode, against the oracle at full size."""
    b, t, hh, ww = 1, 7, 576, 704
    sd = so.make_state_dict(0)
    eng = _engine(dev, sd)
    x = so.make_frames(b, t, hh, ww, 1234)
    eps = so.make_eps(b, t, hh // 4, ww // 4, 99)
    torch.set_num_threads(os.cpu_count() or 1)
    with torch.no_grad():
        z = so.net_down(sd, x, t)
        lr = so.quantize(z[:, :3])
        hr_ref, hf_ref = so.net_up(sd, lr, eps, t)
    out51, lr_u8, lr_q = eng.down(x.to(dev), t)
    torch.testing.assert_close(out51.cpu(), z, rtol=0, atol=2e-4)
    diff = (lr_u8.cpu().int() - so.quantize_u8(z[:, :3]).int()).abs()
    assert diff.max().item() <= 1 and (diff == 0).float().mean().item() >= 0.9999
    hr, hf = eng.up(lr.to(dev), t, eps=eps.to(dev))
    torch.testing.assert_close(hf.cpu(), hf_ref, rtol=0, atol=5e-4)
    assert (hr.cpu() - hr_ref).abs().max().item() <= HR_TOL_FP32
    # per-clip metrics within 0.01 dB / 1e-4 (test_rescaling.py:109-123)
    ya, yb, y0 = so.rgb_to_y(hr.cpu()), so.rgb_to_y(hr_ref), so.rgb_to_y(x)
    p_got, p_ref = np.mean(so.psnr_frames(ya, y0)), np.mean(so.psnr_frames(yb, y0))
    s_got, s_ref = np.mean(so.ssim_frames(ya, y0)), np.mean(so.ssim_frames(yb, y0))
    assert abs(p_got - p_ref) <= 0.01 and abs(s_got - s_ref) <= 1e-4


@pytest.mark.parametrize("b,t,hh,ww", [(2, 7, 96, 160), (1, 1, 20, 28), (1, 3, 36, 44), (1, 9, 16, 16), (1, 1, 4, 4), (3, 5, 8, 12), (1, 2, 484, 16),
                                       (1, 1, 16, 1940)])
def test_bf16_mode_vs_oracle(dev, b, t, hh, ww):
    """bf16 mode (tcgen05 convolutions, bf16 activations, fp32 state): HR within 2e-2, LR within +-1 LSB.
    Shapes cover an odd pixel count (pointwise pseudo-frame split), T=9 (temporal FMA fallback), a single LR pixel, an odd number of
    strip columns (B*T = 15: the fused dense-block kernel's CTA pairs get a dummy column), a tall narrow clip (row ranges of one
    column pair spread over many CTA pairs) and a wide one (485 LR columns = 5 strips, the last 5 columns wide)."""
    sd = so.make_state_dict(0)
    eng = _engine(dev, sd, "bf16")
    x = so.make_frames(b, t, hh, ww, 77)
    eps = so.make_eps(b, t, hh // 4, ww // 4, 5)
    with torch.no_grad():
        z = so.net_down(sd, x, t)
        lr = so.quantize(z[:, :3])
        hr_ref, hf_ref = so.net_up(sd, lr, eps, t)
    out51, lr_u8, _ = eng.down(x.to(dev), t)
    exact, within1, mx = _report_lr(f"bf16 {b}x{t}x{hh}x{ww}", lr_u8.cpu(), so.quantize_u8(z[:, :3]))
    few = lr_u8.numel() < 2000                                            # a handful of codes: fractions are meaningless
    assert mx <= 1 and within1 >= 0.9999 and (few or exact >= LR_EXACT_BF16)      # a systematic 1-LSB offset would fail here
    hr, hf = eng.up(lr.to(dev), t, eps=eps.to(dev))
    assert (hr.cpu() - hr_ref).abs().max().item() <= HR_TOL_BF16


# ------------------------------------------------------------------------------------------------ boundary (8b)
def test_drop_in_module_matches_oracle(dev):
    from selfc_b200 import networks, options
    from selfc_b200.global_var import GlobalVar
    here = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    opt = options.dict_to_nonedict(options.parse(os.path.join(here, "selfc_b200", "configs", "selfc_large_synthetic.yml"),
                                                 is_train=False))
    net = networks.define_G(opt)
    sd = so.make_state_dict(4)
    net.load_state_dict(sd, strict=True)
    net = net.to(dev).eval()
    b, t, hh, ww = 1, 3, 24, 40
    GlobalVar.set_Temporal_LEN(t)
    x = so.make_frames(b, t, hh, ww, 21)
    eps = so.make_eps(b, t, hh // 4, ww // 4, 8)
    with torch.no_grad():
        out, loss_c = net(x=x.to(dev))
        assert out.shape == (b * t, 51, hh // 4, ww // 4) and float(loss_c) == 0.0
        z = so.net_down(sd, x, t)
        torch.testing.assert_close(out.cpu(), z, rtol=0, atol=1e-4)
        lr = so.quantize(z[:, :3])
        net.inject_eps(eps.to(dev))
        hr, hf = net(x=lr.to(dev), rev=True)
        hr_ref, hf_ref = so.net_up(sd, lr, eps, t)
        torch.testing.assert_close(hf.cpu(), hf_ref, rtol=0, atol=5e-4)
        torch.testing.assert_close(hr.cpu(), hr_ref, rtol=0, atol=HR_TOL_FP32)
        assert net.stp_net.gmm_v.shape == (b, 48, t, hh // 4, ww // 4)
    with pytest.raises(RuntimeError):
        net.cpu()(x=x)                      # no CPU fallback


def test_errors_are_loud(dev):
    from selfc_b200.engine import Engine
    eng = Engine(dev, "fp32")
    with pytest.raises(RuntimeError, match="weights not loaded"):
        eng.down(torch.zeros(1, 3, 8, 8, device=dev), 1)
    eng.load_state(so.make_state_dict(0))
    with pytest.raises(ValueError):
        eng.down(torch.zeros(3, 3, 8, 8, device=dev), 2)      # B*T not a multiple of T
    with pytest.raises(ValueError):
        eng.down(torch.zeros(2, 3, 10, 8, device=dev), 2)     # H not a multiple of 4


# ------------------------------------------------------------------------------------------------ full-size properties
def test_1080p_properties(dev):
    """BASELINE.json configs[2] size (one 7-frame 1080p GOP): properties that need no oracle run."""
    b, t, hh, ww = 1, 7, 1080, 1920
    sd = so.make_state_dict(0)
    eng = _engine(dev, sd)
    x = so.make_frames(b, t, hh, ww, 4321).to(dev)
    out51, lr_u8, lr_q = eng.down(x, t)
    # LR codes are exactly the quantisation of the returned latent
    assert torch.equal(lr_u8, torch.round(out51[:, :3].clamp(0, 1) * 255.0).to(torch.uint8))
    # determinism + seed/offset sensitivity + batch invariance of the noise stream
    hr1, _ = eng.up(lr_q, t, seed=7, offset=0)
    hr2, _ = eng.up(lr_q, t, seed=7, offset=0)
    hr3, _ = eng.up(lr_q, t, seed=7, offset=1)
    assert torch.equal(hr1, hr2) and not torch.equal(hr1, hr3)
    assert torch.isfinite(hr1).all()
    # the first frame's top-left corner depends only on its own GOP: a 2-GOP batch reproduces it
    x2 = torch.cat([x, x.flip(0)], 0)
    _, lr2_u8, lr2_q = eng.down(x2, t, want_out51=False)
    assert torch.equal(lr2_u8[:t], lr_u8)


@pytest.mark.parametrize("hh,ww", [(1080, 1920), (2160, 3840)])
def test_full_size_bf16_mode_agrees_with_fp32_mode(dev, hh, ww):
    """BASELINE.json configs[2] / configs[4] sizes (one 7-frame 1080p / 4K GOP): the tcgen05 bf16 path against the fp32-FMA
    path (itself oracle-checked at the sizes the oracle finishes) on the same frames, LR codes and noise stream --
    LR within 1 LSB on >= 99.99 % of the pixels, HR max-abs <= 2e-2 (north star), plus determinism."""
    b, t = 1, 7
    sd = so.make_state_dict(0)
    x = so.make_frames(b, t, hh, ww, 99).to(dev)
    e32 = _engine(dev, sd, "fp32")
    _, lr32_u8, lr32_q = e32.down(x, t, want_out51=False)
    hr32, _ = e32.up(lr32_q, t, seed=5, offset=2, want_hf=False)
    del e32
    torch.cuda.empty_cache()
    e16 = _engine(dev, sd, "bf16")
    _, lr16_u8, _ = e16.down(x, t, want_out51=False)
    d = (lr16_u8.int() - lr32_u8.int()).abs()
    assert d.max().item() <= 1 and (d <= 1).float().mean().item() >= 0.9999   # north star: within +-1 LSB on >= 99.99 %
    hr16, _ = e16.up(lr32_q, t, seed=5, offset=2, want_hf=False)
    hr16b, _ = e16.up(lr32_q, t, seed=5, offset=2, want_hf=False)
    assert torch.equal(hr16, hr16b) and torch.isfinite(hr16).all()
    assert (hr16 - hr32).abs().max().item() <= 2e-2


# ------------------------------------------------------------------------------------------------ benchmark size vs the ORACLE
def _clip_metrics(hr, ref_frames):
    ya, y0 = so.rgb_to_y(hr), so.rgb_to_y(ref_frames)
    return float(np.mean(so.psnr_frames(ya, y0))), float(np.mean(so.ssim_frames(ya, y0)))


@pytest.fixture(scope="module")
def oracle_1080p():
    """The oracle on TWO different 7-frame 1080p GOPs (BASELINE.json configs[2] size; ~20 s of host time each): latent, LR
    codes, HR and the per-clip Y-PSNR / Y-SSIM of test_rescaling.py:109-123.  Weights: the real checkpoint when present."""
    from conftest import test_weights
    t, hh, ww = 7, 1080, 1920
    sd, desc = test_weights(0)
    torch.set_num_threads(os.cpu_count() or 1)
    clips = []
    for c in range(2):
        x = so.make_frames(1, t, hh, ww, 4321 + c)
        eps = so.make_eps(1, t, hh // 4, ww // 4, 17 + c)
        with torch.no_grad():
            z = so.net_down(sd, x, t)
            lr = so.quantize(z[:, :3])
            hr, _ = so.net_up(sd, lr, eps, t)
        clips.append({"x": x, "eps": eps, "lr_u8": so.quantize_u8(z[:, :3]), "lr": lr, "hr": hr, "metrics": _clip_metrics(hr, x)})
        del z
    print(f"[oracle_1080p] weights: {desc}")
    return {"sd": sd, "t": t, "clips": clips}


def _report_lr(tag, got_u8, ref_u8):
    d = (got_u8.int() - ref_u8.int()).abs()
    exact, within1 = (d == 0).float().mean().item(), (d <= 1).float().mean().item()
    print(f"[{tag}] LR codes: exact match {exact * 100:.4f} %, within 1 LSB {within1 * 100:.4f} %, max |diff| {d.max().item()}, "
          f"mean signed diff {(got_u8.float() - ref_u8.float()).mean().item():+.2e} LSB")
    return exact, within1, int(d.max().item())


def test_1080p_gop_fp32_mode_vs_oracle(dev, oracle_1080p):
    """fp32 mode at the benchmark's own size, directly against the oracle: LR exact on >= 99.99 %, HR <= 1e-3, clip metrics."""
    o, c = oracle_1080p, oracle_1080p["clips"][0]
    eng = _engine(dev, o["sd"], "fp32")
    _, lr_u8, _ = eng.down(c["x"].to(dev), o["t"], want_out51=False)
    exact, within1, mx = _report_lr("1080p fp32", lr_u8.cpu(), c["lr_u8"])
    assert mx <= 1 and exact >= 0.9999
    hr, _ = eng.up(c["lr"].to(dev), o["t"], eps=c["eps"].to(dev), want_hf=False)
    err = (hr.cpu() - c["hr"]).abs().max().item()
    p, s = _clip_metrics(hr.cpu(), c["x"])
    print(f"[1080p fp32] HR max |diff| {err:.3e}; Y-PSNR {p:.4f} dB (oracle {c['metrics'][0]:.4f}), Y-SSIM {s:.6f} (oracle {c['metrics'][1]:.6f})")
    assert err <= HR_TOL_FP32
    assert abs(p - c["metrics"][0]) <= 0.01 and abs(s - c["metrics"][1]) <= 1e-4


def test_1080p_two_gops_bf16_mode_vs_oracle(dev, oracle_1080p):
    """bf16 mode (the mode the headline number is quoted in) at the benchmark's size with B = 2, directly against the oracle:
    LR within 1 LSB on >= 99.99 % of the pixels (exact-match fraction and signed bias reported and bounded), HR <= 2e-2, and the
    north star's per-clip gate (Y-PSNR within 0.01 dB, Y-SSIM within 1e-4) for each clip."""
    o = oracle_1080p
    t = o["t"]
    eng = _engine(dev, o["sd"], "bf16")
    x = torch.cat([c["x"] for c in o["clips"]], 0).to(dev)
    ref_u8 = torch.cat([c["lr_u8"] for c in o["clips"]], 0)
    _, lr_u8, _ = eng.down(x, t, want_out51=False)
    exact, within1, mx = _report_lr("1080p bf16 B=2", lr_u8.cpu(), ref_u8)
    # the north star's gate: within 1 LSB on >= 99.99 % of the pixels (over 5.4 M codes a handful land 2 LSB away: measured
    # 100.0000 % within 1, max 2); anything further off than 2 would be a bug, not rounding
    assert within1 >= 0.9999 and mx <= 2
    assert exact >= LR_EXACT_BF16, "a systematic 1-LSB offset would show up as a low exact-match fraction"
    bias = (lr_u8.cpu().float() - ref_u8.float()).mean().item()
    assert abs(bias) <= LR_BIAS_BF16, f"LR codes are biased by {bias} LSB against the oracle"
    lr = torch.cat([c["lr"] for c in o["clips"]], 0).to(dev)
    eps = torch.cat([c["eps"] for c in o["clips"]], 0).to(dev)
    hr, _ = eng.up(lr, t, eps=eps, want_hf=False)
    hr = hr.cpu()
    for i, c in enumerate(o["clips"]):
        got = hr[i * t:(i + 1) * t]
        err = (got - c["hr"]).abs().max().item()
        p, s = _clip_metrics(got, c["x"])
        print(f"[1080p bf16 clip {i}] HR max |diff| {err:.3e}; Y-PSNR {p:.4f} dB (oracle {c['metrics'][0]:.4f}), "
              f"Y-SSIM {s:.6f} (oracle {c['metrics'][1]:.6f})")
        assert err <= HR_TOL_BF16
        assert abs(p - c["metrics"][0]) <= 0.01 and abs(s - c["metrics"][1]) <= 1e-4


def test_vid4_shape_bf16_mode_metrics_gate(dev):
    """configs[1] shape (7 x 576 x 704) in bf16 mode: the per-clip Y-PSNR / Y-SSIM gate of the north star, plus LR / HR gates."""
    from conftest import test_weights
    b, t, hh, ww = 1, 7, 576, 704
    sd, _ = test_weights(0)
    eng = _engine(dev, sd, "bf16")
    x = so.make_frames(b, t, hh, ww, 1234)
    eps = so.make_eps(b, t, hh // 4, ww // 4, 99)
    torch.set_num_threads(os.cpu_count() or 1)
    with torch.no_grad():
        z = so.net_down(sd, x, t)
        lr = so.quantize(z[:, :3])
        hr_ref, _ = so.net_up(sd, lr, eps, t)
    _, lr_u8, _ = eng.down(x.to(dev), t, want_out51=False)
    exact, within1, mx = _report_lr("vid4 bf16", lr_u8.cpu(), so.quantize_u8(z[:, :3]))
    assert mx <= 1 and within1 >= 0.9999 and exact >= LR_EXACT_BF16
    hr, _ = eng.up(lr.to(dev), t, eps=eps.to(dev), want_hf=False)
    assert (hr.cpu() - hr_ref).abs().max().item() <= HR_TOL_BF16
    p_got, s_got = _clip_metrics(hr.cpu(), x)
    p_ref, s_ref = _clip_metrics(hr_ref, x)
    print(f"[vid4 bf16] Y-PSNR {p_got:.4f} dB (oracle {p_ref:.4f}), Y-SSIM {s_got:.6f} (oracle {s_ref:.6f})")
    assert abs(p_got - p_ref) <= 0.01 and abs(s_got - s_ref) <= 1e-4


# ------------------------------------------------------------------------------------------------ BF16X3 mode (x1: configs[1] on tensor cores)
# The numerics-gate configuration on the tcgen05 kernels: (hi, lo) bf16 pairs, three MMAs per product, fp32 accumulation.  The
# north star's fp32 gate applies -- HR <= 1e-3, LR exact on >= 99.99 %, clip metrics within 0.01 dB / 1e-4 -- and the tests hold it
# to a tighter bound (measured 2.3e-5 at Vid4 size) so that a lost partial product (2^-9 ~ 2e-3 relative) cannot hide.
HR_TOL_X3 = 2e-4


@pytest.mark.parametrize("prefix,cin,k", [("operations.5.G", 3, 0), ("operations.2.F", 48, 0), ("operations.2.F", 48, 3),
                                          ("operations.5.H", 3, 1), ("stp_net.local_m2", 64, 3), ("stp_net.local_m1", 3, 2)])
@pytest.mark.parametrize("shape", [(1, 2, 13, 21), (2, 3, 40, 70), (1, 1, 1, 1), (1, 1, 9, 61)])
def test_x3_conv3x3_vs_fp64(dev, prefix, cin, k, shape):
    """One (1,3,3) conv through the (hi, lo) form of the tcgen05 kernel against the same conv in fp64 on the UNROUNDED operands."""
    import torch.nn.functional as F
    sd = so.make_state_dict(9)
    eng = _engine(dev, sd, "bf16x3")
    b, t, h, w = shape
    x = torch.randn(b * t, cin + 32 * k, h, w, generator=torch.Generator().manual_seed(cin + k)) * 0.7
    wgt, bias = sd[f"{prefix}.conv{k + 1}.weight"], sd[f"{prefix}.conv{k + 1}.bias"]
    ref = F.leaky_relu(F.conv2d(x.double(), wgt[:, :, 0].double(), bias.double(), padding=1), 0.2).float()
    got = eng.conv3x3(prefix, k, x.to(dev), t).cpu()
    torch.testing.assert_close(got, ref, rtol=2e-5, atol=5e-5)


@pytest.mark.parametrize("prefix,cin", [("operations.1.F", 48), ("operations.3.G", 3), ("stp_net.local_m2", 64), ("stp_net.local_m1", 3)])
@pytest.mark.parametrize("t,h,w", [(1, 9, 14), (3, 13, 21), (7, 24, 40), (9, 8, 8), (16, 5, 7)])
def test_x3_dense_block_vs_oracle(dev, prefix, cin, t, h, w):
    sd = so.make_state_dict(5)
    eng = _engine(dev, sd, "bf16x3")
    b = 2
    x = torch.randn(b * t, cin, h, w, generator=torch.Generator().manual_seed(7)) * 0.5
    with torch.no_grad():
        ref = so.d2dt(sd, prefix, x, t)
    got = eng.d2dt(prefix, x.to(dev), t).cpu()
    torch.testing.assert_close(got, ref, rtol=2e-5, atol=3e-5)


@pytest.mark.parametrize("h,w,t", [(10, 18, 2), (45, 67, 7), (40, 52, 12)])
def test_x3_global_agg_vs_oracle(dev, h, w, t):
    sd = so.make_state_dict(6, gain=2.0)
    eng = _engine(dev, sd, "bf16x3")
    b = 2
    x = torch.randn(b * t, 64, h, w, generator=torch.Generator().manual_seed(h))
    with torch.no_grad():
        ref = so.global_agg(sd, "stp_net.global_m2", x, t)
        wref = so.global_agg_weights(sd, "stp_net.global_m2", x, t)
    got, wmat = eng.global_agg("stp_net.global_m2", x.to(dev), t)
    torch.testing.assert_close(wmat.cpu(), wref, rtol=0, atol=1e-5)
    torch.testing.assert_close(got.cpu(), ref, rtol=5e-5, atol=2e-4)


@pytest.mark.parametrize("b,t,hh,ww", [(2, 7, 96, 160), (1, 1, 20, 28), (1, 3, 36, 44), (1, 9, 16, 16), (1, 1, 4, 4), (3, 5, 8, 12), (1, 2, 484, 16),
                                       (1, 1, 16, 1940), (1, 12, 24, 36)])
def test_x3_mode_vs_oracle(dev, b, t, hh, ww):
    """Whole network in BF16X3 mode on the shapes the bf16 mode is tested on (odd pixel counts, a single LR pixel, T up to 12, tall
    and wide clips): latent and HR far inside the fp32 gate, LR codes exact but for rounding-boundary cases."""
    sd = so.make_state_dict(0)
    eng = _engine(dev, sd, "bf16x3")
    x = so.make_frames(b, t, hh, ww, 77)
    eps = so.make_eps(b, t, hh // 4, ww // 4, 5)
    with torch.no_grad():
        z = so.net_down(sd, x, t)
        lr = so.quantize(z[:, :3])
        hr_ref, hf_ref = so.net_up(sd, lr, eps, t)
    out51, lr_u8, _ = eng.down(x.to(dev), t)
    assert (out51.cpu() - z).abs().max().item() <= HR_TOL_X3
    exact, within1, mx = _report_lr(f"bf16x3 {b}x{t}x{hh}x{ww}", lr_u8.cpu(), so.quantize_u8(z[:, :3]))
    # codes that differ sit on a rounding boundary (latent within 2e-4 of it): a handful per clip, whatever its size
    assert mx <= 1 and (exact >= 0.999 or round((1.0 - exact) * lr_u8.numel()) <= 3)
    hr, hf = eng.up(lr.to(dev), t, eps=eps.to(dev))
    torch.testing.assert_close(hf.cpu(), hf_ref, rtol=0, atol=5e-4)
    assert (hr.cpu() - hr_ref).abs().max().item() <= HR_TOL_X3
    # the Philox stream and the 8-bit interface go through the same kernels as in the other modes
    hr_a, _ = eng.up(lr.to(dev), t, seed=3, offset=1, want_hf=False)
    hr_b, _ = eng.up(lr.to(dev), t, seed=3, offset=1, want_hf=False)
    assert torch.equal(hr_a, hr_b) and torch.isfinite(hr_a).all()


def test_x3_vid4_shape_batch_vs_oracle(dev):
    """BASELINE.json configs[1]: batched 7-frame 576x704 clips (B = 2 here, two different clips) on tensor cores against the oracle at
    full size -- the fp32 gate (HR <= 1e-3; held to 2e-4), LR exact on >= 99.9 % and within 1 LSB everywhere, per-clip Y-PSNR /
    Y-SSIM within 0.01 dB / 1e-4."""
    from conftest import test_weights
    b, t, hh, ww = 2, 7, 576, 704
    sd, _ = test_weights(0)
    eng = _engine(dev, sd, "bf16x3")
    x = torch.cat([so.make_frames(1, t, hh, ww, 1234), so.make_frames(1, t, hh, ww, 4321)], 0)
    eps = so.make_eps(b, t, hh // 4, ww // 4, 99)
    torch.set_num_threads(os.cpu_count() or 1)
    with torch.no_grad():
        z = so.net_down(sd, x, t)
        lr = so.quantize(z[:, :3])
        hr_ref, _ = so.net_up(sd, lr, eps, t)
    _, lr_u8, _ = eng.down(x.to(dev), t, want_out51=False)
    exact, within1, mx = _report_lr("vid4 bf16x3 B=2", lr_u8.cpu(), so.quantize_u8(z[:, :3]))
    assert mx <= 1 and exact >= 0.999
    hr, _ = eng.up(lr.to(dev), t, eps=eps.to(dev), want_hf=False)
    hr = hr.cpu()
    err = (hr - hr_ref).abs().max().item()
    print(f"[vid4 bf16x3 B=2] HR max |diff| {err:.3e}")
    assert err <= HR_TOL_X3
    for i in range(b):
        sl = slice(i * t, (i + 1) * t)
        p_got, s_got = _clip_metrics(hr[sl], x[sl])
        p_ref, s_ref = _clip_metrics(hr_ref[sl], x[sl])
        print(f"[vid4 bf16x3 clip {i}] Y-PSNR {p_got:.4f} dB (oracle {p_ref:.4f}), Y-SSIM {s_got:.6f} (oracle {s_ref:.6f})")
        assert abs(p_got - p_ref) <= 0.01 and abs(s_got - s_ref) <= 1e-4


def test_x3_1080p_gop_vs_oracle(dev, oracle_1080p):
    """BF16X3 mode at the benchmark's size against the oracle: the fp32 gate on tensor cores."""
    o, c = oracle_1080p, oracle_1080p["clips"][1]
    eng = _engine(dev, o["sd"], "bf16x3")
    _, lr_u8, _ = eng.down(c["x"].to(dev), o["t"], want_out51=False)
    exact, within1, mx = _report_lr("1080p bf16x3", lr_u8.cpu(), c["lr_u8"])
    assert mx <= 1 and exact >= 0.999
    hr, _ = eng.up(c["lr"].to(dev), o["t"], eps=c["eps"].to(dev), want_hf=False)
    err = (hr.cpu() - c["hr"]).abs().max().item()
    p, s = _clip_metrics(hr.cpu(), c["x"])
    print(f"[1080p bf16x3] HR max |diff| {err:.3e}; Y-PSNR {p:.4f} dB (oracle {c['metrics'][0]:.4f}), Y-SSIM {s:.6f} (oracle {c['metrics'][1]:.6f})")
    assert err <= HR_TOL_X3
    assert abs(p - c["metrics"][0]) <= 0.01 and abs(s - c["metrics"][1]) <= 1e-4


# ------------------------------------------------------------------------------------------------ tcgen05 conv (bf16 mode)
def _bf16r(t):
    return t.to(torch.bfloat16).to(torch.float32)


@pytest.mark.parametrize("prefix,cin,k", [("operations.2.F", 48, 0), ("operations.2.F", 48, 3), ("operations.5.G", 3, 0),
                                          ("operations.5.H", 3, 1), ("stp_net.local_m2", 64, 3), ("stp_net.local_m1", 3, 2)])
@pytest.mark.parametrize("shape", [(1, 2, 13, 21), (2, 3, 40, 70)])
def test_conv3x3_tc_vs_oracle(dev, prefix, cin, k, shape):
    """One (1,3,3) conv through the tcgen05 kernel vs conv3d on bf16-rounded operands (fp32 accumulate)."""
    import torch.nn.functional as F
    sd = so.make_state_dict(9)
    eng = _engine(dev, sd, "bf16")
    b, t, h, w = shape
    x = torch.randn(b * t, cin + 32 * k, h, w, generator=torch.Generator().manual_seed(cin + k)) * 0.7
    wgt, bias = sd[f"{prefix}.conv{k + 1}.weight"], sd[f"{prefix}.conv{k + 1}.bias"]
    ref = F.leaky_relu(F.conv2d(_bf16r(x), _bf16r(wgt)[:, :, 0], bias, padding=1), 0.2)
    got = eng.conv3x3(prefix, k, x.to(dev), t).cpu()
    torch.testing.assert_close(got, ref, rtol=1e-2, atol=3e-3)


@pytest.mark.parametrize("prefix,cin", [("operations.1.F", 48), ("operations.3.G", 3), ("stp_net.local_m2", 64)])
@pytest.mark.parametrize("t,h,w", [(1, 9, 14), (3, 13, 21), (7, 24, 40), (9, 8, 8)])
def test_dense_block_bf16_vs_oracle(dev, prefix, cin, t, h, w):
    """Whole D2DTInput in bf16 mode (tcgen05 conv1-4 + tcgen05 temporal conv5; T=9 exceeds TMEM and takes the FMA kernel)."""
    sd = so.make_state_dict(5)
    eng = _engine(dev, sd, "bf16")
    b = 2
    x = _bf16r(torch.randn(b * t, cin, h, w, generator=torch.Generator().manual_seed(7)) * 0.5)
    with torch.no_grad():
        ref = so.d2dt(sd, prefix, x, t)
    got = eng.d2dt(prefix, x.to(dev), t).cpu()
    torch.testing.assert_close(got, ref, rtol=2e-2, atol=6e-3)


@pytest.mark.parametrize("h,w,t", [(10, 18, 2), (45, 67, 7)])
def test_global_agg_bf16_vs_oracle(dev, h, w, t):
    sd = so.make_state_dict(6, gain=2.0)
    eng = _engine(dev, sd, "bf16")
    b = 2
    x = _bf16r(torch.randn(b * t, 64, h, w, generator=torch.Generator().manual_seed(h)))
    with torch.no_grad():
        ref = so.global_agg(sd, "stp_net.global_m2", x, t)
        wref = so.global_agg_weights(sd, "stp_net.global_m2", x, t)
    got, wmat = eng.global_agg("stp_net.global_m2", x.to(dev), t)
    torch.testing.assert_close(wmat.cpu(), wref, rtol=0, atol=1e-5)
    torch.testing.assert_close(got.cpu(), ref, rtol=2e-2, atol=2e-2)


# ------------------------------------------------------------------------------------------------ callers (a12, f1)
def test_gaussian_lr_ref_vs_oracle(dev):
    from selfc_b200 import engine
    x = torch.rand(5, 3, 40, 56, generator=torch.Generator().manual_seed(2))
    got = engine.gaussian_downsample(x.to(dev)).cpu()
    torch.testing.assert_close(got, so.gaussian_downsample(x), rtol=0, atol=2e-6)


def test_model_wrapper_matches_oracle(dev, tmp_path):
    """create_model(opt) -> feed_data -> test -> get_current_visuals, the call sequence of test_rescaling.py:65-110,
    with a checkpoint file in the reference's format (DataParallel 'module.' prefix)."""
    from selfc_b200 import engine, model, options
    from selfc_b200.global_var import GlobalVar
    here = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sd = so.make_state_dict(12)
    ckpt = tmp_path / "selfc_large_synthetic.pth"
    torch.save({"module." + k: v for k, v in sd.items()}, ckpt)
    opt = options.parse(os.path.join(here, "selfc_b200", "configs", "selfc_large_synthetic.yml"), is_train=False)
    opt["path"]["pretrain_model_G"] = str(ckpt)
    opt = options.dict_to_nonedict(opt)
    b, t, hh, ww = 2, 7, 32, 48
    GlobalVar.set_Temporal_LEN(t)
    m = model.create_model(opt)
    frames = so.make_frames(b, t, hh, ww, 31)                            # [b*t,3,H,W]
    gt = frames.reshape(b, t, 3, hh, ww).transpose(1, 2).contiguous()    # dataset layout [B,3,T,H,W]
    m.netG.module.set_noise(seed=5, offset=9)
    assert m.feed_data({"GT": gt}) == t
    m.test()
    vis = m.get_current_visuals()
    assert set(vis) == {"SR", "LR_ref", "LR", "GT", "forw_H"}
    eps = engine.export_eps(b, t, hh // 4, ww // 4, seed=5, offset=9, device=dev).cpu()
    with torch.no_grad():
        z = so.net_down(sd, frames, t)
        lr = so.quantize(z[:, :3])
        hr_ref, _ = so.net_up(sd, lr, eps, t)
    assert torch.equal(vis["GT"].cpu(), frames)
    torch.testing.assert_close(vis["LR_ref"].cpu(), so.gaussian_downsample(frames), rtol=0, atol=2e-6)
    d = (torch.round(vis["LR"].cpu() * 255) - torch.round(lr * 255)).abs()
    assert d.max().item() <= 1 and (d == 0).float().mean().item() >= 0.9999
    torch.testing.assert_close(vis["forw_H"].cpu(), z[:, 3:], rtol=0, atol=1e-4)
    # the wrapper quantised its own LR; feed the oracle the same codes for the HR comparison
    hr_ref2, _ = so.net_up(sd, vis["LR"].cpu(), eps, t)
    torch.testing.assert_close(vis["SR"].cpu(), hr_ref2, rtol=0, atol=HR_TOL_FP32)


def test_rescale_host_pipeline_matches_per_gop_calls(dev):
    """The overlapped host-buffer API gives bit-identical results to plain per-GOP calls (incl. the padded tail GOP)."""
    sd = so.make_state_dict(0)
    eng = _engine(dev, sd, "bf16")
    n, hh, ww = 16, 32, 48            # 2 full GOPs + a 2-frame tail
    frames = so.make_frames(1, n, hh, ww, 3)
    host_in = frames.pin_memory()
    lr_h = torch.empty(n, 3, hh // 4, ww // 4, dtype=torch.uint8).pin_memory()
    hr_h = torch.empty(n, 3, hh, ww).pin_memory()
    eng.rescale_host(host_in, lr_h, hr_h, 7, seed=3, offset0=10)
    from selfc_b200.sharding import gop_indices
    for i, (ids, real) in enumerate(gop_indices(n, 7)):
        lr_u8, rec = eng.rescale(frames[ids].to(dev), 7, seed=3, offset=10 + i)
        assert torch.equal(lr_h[ids[0]:ids[0] + real], lr_u8[:real].cpu())
        assert torch.equal(hr_h[ids[0]:ids[0] + real], rec[:real].cpu())


# ------------------------------------------------------------------------------------------------ f2: 8-bit frames at the boundary
def test_u8_frame_conversions_bit_exact(dev, golden_dir):
    """Stand-alone ingest/egress kernels against the reference's helpers (fixture) and the oracle (odd sizes, ties, out of range)."""
    eng = _engine(dev, so.make_state_dict(0))
    g = np.load(os.path.join(golden_dir, "u8.npz"))
    assert np.array_equal(eng.frames_from_u8(torch.from_numpy(g["img"]).to(dev)).cpu().numpy(), g["from_ref"])
    assert np.array_equal(eng.frames_to_u8(_t(g["y"]).to(dev)).cpu().numpy(), g["to_ref"])
    rng = np.random.default_rng(4)
    img = rng.integers(0, 256, size=(3, 37, 53, 3), dtype=np.uint8)
    assert np.array_equal(eng.frames_from_u8(torch.from_numpy(img).to(dev)).cpu().numpy(), so.frames_from_u8(img).numpy())
    y = torch.randn(3, 3, 37, 53, generator=torch.Generator().manual_seed(5)) * 0.5 + 0.5
    y.view(-1)[:512] = (torch.arange(512, dtype=torch.float32) * 0.5) / 255.0        # every half code: all the ties
    assert np.array_equal(eng.frames_to_u8(y.to(dev)).cpu().numpy(), so.frames_to_u8(y))
    assert eng.frames_to_u8(torch.zeros(0, 3, 8, 8, device=dev)).shape == (0, 8, 8, 3)


@pytest.mark.parametrize("mode", ["fp32", "bf16", "bf16x3"])
def test_u8_path_equals_fp32_interface_on_converted_frames(dev, mode):
    """The fused 8-bit entry points are the fp32 entry points composed with the CPU conversions, bit for bit:
    down_u8(img) == to_u8(down(from_u8(img))), up_u8(lr_img) == to_u8(up(from_u8(lr_img))), and the one-call rescale_u8."""
    sd = so.make_state_dict(0)
    eng = _engine(dev, sd, mode)
    b, t, hh, ww = 2, 3, 40, 72
    rng = np.random.default_rng(6)
    base = so.make_frames(b, t, hh, ww, 8)
    img = so.frames_to_u8(base)                                  # a plausible 8-bit clip
    img[0, :2, :8] = rng.integers(0, 256, size=(2, 8, 3), dtype=np.uint8)
    x = so.frames_from_u8(img)
    img_d = torch.from_numpy(img).to(dev)
    _, lr_u8, lr_q = eng.down(x.to(dev), t, want_out51=False)
    lr_img, lr_q2 = eng.down_u8(img_d, t, want_q=True)
    assert torch.equal(lr_q2, lr_q)
    assert np.array_equal(lr_img.cpu().numpy(), so.frames_to_u8(lr_q.cpu()))
    assert np.array_equal(lr_img.cpu().numpy(), np.transpose(lr_u8.cpu().numpy()[:, [2, 1, 0]], (0, 2, 3, 1)))
    eps = so.make_eps(b, t, hh // 4, ww // 4, 15).to(dev)
    hr, _ = eng.up(lr_q, t, eps=eps, want_hf=False)
    hr_img = eng.up_u8(lr_img, t, eps=eps)
    assert np.array_equal(hr_img.cpu().numpy(), so.frames_to_u8(hr.cpu()))
    lr_img2, hr_img2 = eng.rescale_u8(img_d, t, eps=eps)
    assert torch.equal(lr_img2, lr_img) and torch.equal(hr_img2, hr_img)
    # Philox noise path too
    hr_p, _ = eng.up(lr_q, t, seed=5, offset=2, want_hf=False)
    assert np.array_equal(eng.up_u8(lr_img, t, seed=5, offset=2).cpu().numpy(), so.frames_to_u8(hr_p.cpu()))


def test_u8_rescale_vs_oracle(dev):
    """8-bit frames through the whole path against the oracle's rescale_u8 (fp32 mode): LR codes within 1 LSB on >= 99.99 %,
    HR codes within 1 (HR tolerance 1e-3 is < half a code, so only rounding ties can differ)."""
    sd = so.make_state_dict(1)
    eng = _engine(dev, sd)
    b, t, hh, ww = 1, 3, 32, 48
    img = so.frames_to_u8(so.make_frames(b, t, hh, ww, 21))
    eps = so.make_eps(b, t, hh // 4, ww // 4, 22)
    lr_ref, hr_ref = so.rescale_u8(sd, img, eps, t)
    lr_img = eng.down_u8(torch.from_numpy(img).to(dev), t)
    dl = np.abs(lr_img.cpu().numpy().astype(int) - lr_ref.astype(int))
    assert dl.max() <= 1 and (dl != 0).sum() <= max(1, int(1e-4 * dl.size))      # 864 codes: at most one rounding tie
    # up from the ORACLE's LR frames so both sides start from identical codes
    hr_img = eng.up_u8(torch.from_numpy(lr_ref).to(dev), t, eps=eps.to(dev))
    dh = np.abs(hr_img.cpu().numpy().astype(int) - hr_ref.astype(int))
    assert dh.max() <= 1 and (dh == 0).mean() >= 0.99


def test_rescale_host_u8_pipeline_matches_per_gop_calls(dev):
    sd = so.make_state_dict(0)
    eng = _engine(dev, sd, "bf16")
    n, hh, ww = 16, 32, 48
    img = torch.from_numpy(so.frames_to_u8(so.make_frames(1, n, hh, ww, 3)))
    host_in = img.pin_memory()
    lr_h = torch.empty(n, hh // 4, ww // 4, 3, dtype=torch.uint8).pin_memory()
    hr_h = torch.empty(n, hh, ww, 3, dtype=torch.uint8).pin_memory()
    eng.rescale_host_u8(host_in, lr_h, hr_h, 7, seed=3, offset0=10)
    from selfc_b200.sharding import gop_indices
    for i, (ids, real) in enumerate(gop_indices(n, 7)):
        lr_img, rec = eng.rescale_u8(img[ids].to(dev), 7, seed=3, offset=10 + i)
        assert torch.equal(lr_h[ids[0]:ids[0] + real], lr_img[:real].cpu())
        assert torch.equal(hr_h[ids[0]:ids[0] + real], rec[:real].cpu())


def test_module_rescale_u8_matches_forward_chain(dev):
    """SelfCInvNet.rescale_u8 == read_img1-style conversion -> forward -> Quantization -> forward(rev) -> tensor2img, bit for bit."""
    from selfc_b200 import networks, options
    from selfc_b200.global_var import GlobalVar
    here = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    opt = options.dict_to_nonedict(options.parse(os.path.join(here, "selfc_b200", "configs", "selfc_large_synthetic.yml"),
                                                 is_train=False))
    net = networks.define_G(opt)
    sd = so.make_state_dict(2)
    net.load_state_dict(sd, strict=True)
    net = net.to(dev).eval()
    t = 3
    GlobalVar.set_Temporal_LEN(t)
    img = so.frames_to_u8(so.make_frames(1, t, 32, 48, 31))
    net.set_noise(9, 4)
    lr_img, hr_img = net.rescale_u8(torch.from_numpy(img).to(dev))
    net.set_noise(9, 4)
    out, _ = net(so.frames_from_u8(img).to(dev))
    lr = so.quantize(out[:, :3].cpu())
    hr, _ = net(lr.to(dev), rev=True)
    assert np.array_equal(lr_img.cpu().numpy(), so.frames_to_u8(lr))
    assert np.array_equal(hr_img.cpu().numpy(), so.frames_to_u8(hr.cpu()))


def test_u8_errors_are_loud(dev):
    eng = _engine(dev, so.make_state_dict(0))
    with pytest.raises(ValueError):
        eng.down_u8(torch.zeros(3, 30, 48, 3, dtype=torch.uint8, device=dev), 3)      # H not a multiple of 4
    with pytest.raises(ValueError):
        eng.down_u8(torch.zeros(3, 3, 32, 48, dtype=torch.uint8, device=dev), 3)      # NCHW instead of cv2 layout
    with pytest.raises(ValueError):
        eng.up_u8(torch.zeros(4, 8, 12, 3, dtype=torch.uint8, device=dev), 3)         # not a multiple of T
    with pytest.raises(RuntimeError):
        eng.down_u8(torch.zeros(3, 32, 48, 3, dtype=torch.uint8), 3)                  # host tensor: no CPU path


# ------------------------------------------------------------------------------------------------ opt-in kernel variants
_VARIANT_SNIPPET = r"""
import sys, torch
sys.path.insert(0, %r)
from oracle import selfc_oracle as so
from selfc_b200.engine import Engine
dev = torch.device("cuda", 0)
eng = Engine(dev, "bf16"); eng.load_state(so.make_state_dict(0))
x = so.make_frames(1, 3, 72, 136, 5).to(dev)          # 3 x 5 tiles per frame: odd tile counts, partial tiles
lr, rec = eng.rescale(x, 3, seed=7, offset=1)
torch.save({"lr": lr.cpu(), "rec": rec.cpu()}, sys.argv[1])
"""


@pytest.mark.parametrize("env", [{"SELFC_TC3_PAIR": "0"}, {"SELFC_TC3_P2": "0"}, {"SELFC_TC3_PAIR": "0", "SELFC_TC3_P2": "0"},
                                 {"SELFC_ZIGZAG": "0"}, {"SELFC_DB_FUSED": "0"},
                                 {"SELFC_DB_FUSED": "0", "SELFC_TC3_PAIR": "0", "SELFC_TC3_P2": "0"}])
def test_conv3x3_variants_are_bit_identical(dev, tmp_path, env):
    """CTA pairs (cta_group::2), position-pair TMA rows and the tile-sweep direction change how conv3x3 is scheduled and fed, not
    what it computes: the bf16 path must give bit-identical LR codes and HR frames with each of them switched off (every variant
    in its own process: the knobs are read once)."""
    import subprocess
    import sys
    here = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    outs = []
    for i, extra in enumerate(({}, env)):
        out = str(tmp_path / f"v{i}.pt")
        e = dict(os.environ)
        for k in ("SELFC_TC3_PAIR", "SELFC_TC3_P2", "SELFC_ZIGZAG", "SELFC_DB_FUSED"):
            e.pop(k, None)
        e["SELFC_F5"] = "0"      # conv5's taps inside the fused launch re-associate an fp32 sum: compared with a tolerance elsewhere
        e.update(extra)
        r = subprocess.run([sys.executable, "-c", _VARIANT_SNIPPET % here, out], env=e, capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stderr[-2000:]
        outs.append(torch.load(out))
    assert torch.equal(outs[0]["lr"], outs[1]["lr"])
    assert torch.equal(outs[0]["rec"], outs[1]["rec"])


_PACK_SNIPPET = r"""
import sys, torch
sys.path.insert(0, %r)
from oracle import selfc_oracle as so
from selfc_b200.engine import Engine
dev = torch.device("cuda", 0)
x = so.make_frames(1, 3, 40, 72, 5).to(dev)
out = {}
for mode in ("bf16", "bf16x3", "fp32"):
    eng = Engine(dev, mode); eng.load_state(so.make_state_dict(2))
    lr, rec = eng.rescale(x, 3, seed=7, offset=1)
    out[mode + ".lr"], out[mode + ".rec"] = lr.cpu(), rec.cpu()
    eng.load_state(so.make_state_dict(4))          # a second load into the same context: other values through the same job tables
    lr, rec = eng.rescale(x, 3, seed=7, offset=1)
    out[mode + ".lr2"], out[mode + ".rec2"] = lr.cpu(), rec.cpu()
torch.save(out, sys.argv[1])
"""


def test_batched_weight_packing_is_bit_identical(dev, tmp_path):
    """selfc_ctx_load_weights records its ~300 pack jobs and runs each kind in one launch (csrc/pack_batch.h); SELFC_PACK_BATCH=0 packs
    tensor by tensor as before.  Same images either way: all three modes must give bit-identical LR codes and HR frames, also after a
    second load into the same context (the cached job tables are reused; the weights differ)."""
    import subprocess
    import sys
    here = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    outs = []
    for i, v in enumerate(("1", "0")):
        out = str(tmp_path / f"p{i}.pt")
        e = dict(os.environ)
        e["SELFC_PACK_BATCH"] = v
        r = subprocess.run([sys.executable, "-c", _PACK_SNIPPET % here, out], env=e, capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stderr[-2000:]
        outs.append(torch.load(out))
    assert set(outs[0]) == set(outs[1]) and len(outs[0]) == 12
    for k in outs[0]:
        assert torch.equal(outs[0][k], outs[1][k]), k
    assert not torch.equal(outs[0]["bf16.rec"], outs[0]["bf16.rec2"])      # the second load did change the weights


_FUSED_SNIPPET = r"""
import sys, torch
sys.path.insert(0, %r)
from oracle import selfc_oracle as so
from selfc_b200.engine import Engine
dev = torch.device("cuda", 0)
b, t, h, w = [int(v) for v in sys.argv[2:6]]
eng = Engine(dev, "bf16"); eng.load_state(so.make_state_dict(3))
out = {}
for prefix, cin in (("operations.4.F", 48), ("operations.7.G", 3), ("stp_net.local_m1", 3), ("stp_net.local_m2", 64)):
    x = (torch.randn(b * t, cin, h, w, generator=torch.Generator().manual_seed(cin + h)) * 0.5).to(dev)
    out[prefix] = eng.d2dt(prefix, x, t).cpu()                              # conv1..4 (+ conv5) of one dense block
    xc = (torch.randn(b * t, cin + 96, h, w, generator=torch.Generator().manual_seed(cin + w)) * 0.5).to(dev)
    out[prefix + ".conv4"] = eng.conv3x3(prefix, 3, xc, t).cpu()            # single layer: always the layer-by-layer kernel
frames = so.make_frames(b, t, 4 * h, 4 * w, 5).to(dev)
lr, rec = eng.rescale(frames, t, seed=7, offset=1)                          # the whole path: dual G+H launches, both directions
out["lr"], out["rec"] = lr.cpu(), rec.cpu()
torch.save(out, sys.argv[1])
"""


@pytest.mark.parametrize("b,t,h,w", [(1, 2, 9, 14), (1, 3, 37, 121), (2, 7, 64, 112), (1, 7, 135, 250), (1, 1, 270, 480)])
def test_fused_dense_block_is_bit_identical_to_layer_by_layer(dev, tmp_path, b, t, h, w):
    """dense_fused_kernel (conv1..4 of a block in one launch, growth channels in tensor memory, strips of 120 columns walked row
    by row) against four conv3x3_tc3_kernel launches: same MMAs in the same K order and the same epilogue arithmetic, so the
    dense blocks' outputs, the LR codes and the HR frames must be IDENTICAL.  Shapes: one partial strip; a strip boundary one
    column wide (121); Vimeo LR shape with B = 2; several strips, odd strip-column count and row ranges that cross column
    boundaries; one full 1080p LR frame (4 exact strips, an odd number of strip columns per pair walk)."""
    import subprocess
    import sys
    here = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    outs = []
    # default (fused blocks + conv5's taps of the F blocks inside the fused launch), SELFC_F5=0 (fused blocks, temporal kernel for
    # every conv5), SELFC_DB_FUSED=0 (layer by layer)
    for i, extra in enumerate(({}, {"SELFC_F5": "0"}, {"SELFC_DB_FUSED": "0"})):
        out = str(tmp_path / f"f{i}.pt")
        e = dict(os.environ)
        for k in ("SELFC_DB_FUSED", "SELFC_F5"):
            e.pop(k, None)
        e.update(extra)
        r = subprocess.run([sys.executable, "-c", _FUSED_SNIPPET % here, out, str(b), str(t), str(h), str(w)], env=e, capture_output=True,
                           text=True, timeout=600)
        assert r.returncode == 0, r.stderr[-2000:]
        outs.append(torch.load(out))
    for k in outs[1]:
        same = torch.equal(outs[1][k], outs[2][k])
        if not same:
            d = (outs[1][k].float() - outs[2][k].float()).abs()
            print(f"[fused vs layer-by-layer] {k}: max |diff| {d.max().item():.3e}, differing {100.0 * (d > 0).float().mean().item():.3f} %")
        assert same, f"{k} differs between the fused and the layer-by-layer dense block"
    # The conv5-taps path sums the three taps' fp32 partial products outside the accumulator: the same numbers up to fp32 summation
    # order (1e-7; bit-identical when T = 1).  In bf16 mode such a perturbation occasionally flips the bf16 rounding of a y1 / y2
    # copy (2^-9 relative), which the following blocks amplify locally to ~1e-3 -- the mode's own noise floor against the oracle
    # (LR codes 89 % exact, HR 5e-3..9e-3) -- so the two paths are compared with that floor (each is within ~1e-2 of the oracle,
    # measured 1.2e-2 apart at worst), not bit for bit.
    for k in outs[0]:
        if k in ("lr", "rec"):
            continue
        assert torch.equal(outs[0][k], outs[1][k]), f"{k}: the dense-block entry points do not use the conv5-taps path"
    dl = (outs[0]["lr"].int() - outs[1]["lr"].int()).abs()
    dr = (outs[0]["rec"] - outs[1]["rec"]).abs().max().item()
    print(f"[conv5 taps in the fused launch vs temporal kernel] LR codes equal {100.0 * (dl == 0).float().mean().item():.4f} %, HR max |diff| {dr:.3e}")
    assert dl.max().item() <= 1 and (dl == 0).float().mean().item() >= 0.95 and dr <= HR_TOL_BF16


# ------------------------------------------------------------------------------------------------ f1: metrics kernels
def test_metrics_kernels_vs_reference_fixture_and_oracle(dev, golden_dir):
    """rgb_to_ycbcr bit-exact; PSNR within 1e-3 dB and SSIM within 1e-5 of the reference's own functions (fixture) and of the
    oracle on a larger odd-sized clip, RGB and fused-luma variants (north star: 0.01 dB / 1e-4)."""
    from selfc_b200 import metrics
    g = np.load(os.path.join(golden_dir, "metrics.npz"))
    a, b = _t(g["a"]), _t(g["b"])
    ya = metrics.rgb_to_ycbcr(a.to(dev))
    assert torch.equal(ya.cpu(), so.rgb_to_y(a))
    np.testing.assert_allclose(ya.cpu().numpy(), g["ya"], rtol=0, atol=1e-6)
    np.testing.assert_allclose(metrics.calculate_psnr(a.to(dev), b.to(dev), to_y=True), g["psnr"], rtol=0, atol=1e-3)
    np.testing.assert_allclose(metrics.calculate_ssim(a.to(dev), b.to(dev), to_y=True), g["ssim"], rtol=0, atol=1e-5)
    np.testing.assert_allclose(metrics.calculate_ssim(ya, metrics.rgb_to_ycbcr(b.to(dev))), g["ssim"], rtol=0, atol=1e-5)
    assert metrics.calculate_psnr(a.to(dev), a.to(dev)) == float("inf")
    gen = torch.Generator().manual_seed(77)
    x = torch.rand(3, 3, 75, 133, generator=gen)
    y = (x + 0.03 * torch.randn(x.shape, generator=gen)).clamp(0, 1)
    np.testing.assert_allclose(metrics.calculate_psnr(x.to(dev), y.to(dev)), so.psnr_frames(x, y), rtol=0, atol=1e-3)
    ref_rgb = [float(np.mean([so.ssim_frames(x[i:i + 1, c:c + 1], y[i:i + 1, c:c + 1])[0] for c in range(3)])) for i in range(3)]
    np.testing.assert_allclose(metrics.calculate_ssim(x.to(dev), y.to(dev)), ref_rgb, rtol=0, atol=1e-5)
    cm = metrics.clip_metrics(x.to(dev), y.to(dev), x[:, :, :32, :40].to(dev), y[:, :, :32, :40].to(dev))
    assert sorted(cm) == ["lr_psnr", "lr_psnr_y", "lr_ssim", "lr_ssim_y", "psnr", "psnr_y", "ssim", "ssim_y"]
    np.testing.assert_allclose(cm["psnr_y"], so.psnr_frames(so.rgb_to_y(x), so.rgb_to_y(y)), rtol=0, atol=1e-3)
    np.testing.assert_allclose(cm["ssim_y"], so.ssim_frames(so.rgb_to_y(x), so.rgb_to_y(y)), rtol=0, atol=1e-5)


# ------------------------------------------------------------------------------------------------ f4: checkpoint / resume
def test_training_state_roundtrip_and_resume(dev, tmp_path):
    """train 2 steps -> save model + .state -> fresh model resumes -> step 3 equals the uninterrupted run's step 3 (losses and
    weights to float-atomics noise); the .state optimizer entry loads into torch.optim.Adam (the reference's optimizer)."""
    from selfc_b200 import options, train_loop
    from selfc_b200.model import create_model
    from selfc_b200.global_var import GlobalVar
    here = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    yml = os.path.join(here, "selfc_b200", "configs", "selfc_large_train_synthetic.yml")

    def make(pretrain=None):
        opt = options.parse(yml, is_train=True)
        opt["path"]["models"] = str(tmp_path / "models")
        opt["path"]["training_state"] = str(tmp_path / "training_state")
        opt["path"]["pretrain_model_G"] = pretrain
        opt["logger"] = {"print_freq": 1, "save_checkpoint_freq": 2}
        opt["train"]["niter"] = 3
        opt = options.dict_to_nonedict(opt)
        torch.manual_seed(5)
        m = create_model(opt)
        m.netG.module.set_noise(11, 0)
        return opt, m

    t = 3
    GlobalVar.set_Temporal_LEN(t)
    ds = train_loop.SyntheticClips(n=3, t=t, size=32, seed=2)
    loader = [{"GT": ds[i]["GT"][None]} for i in range(3)]
    opt, m_full = make()
    sd0 = {k: v.detach().clone() for k, v in m_full.netG.module.state_dict().items()}
    assert train_loop.train(opt, m_full, loader, total_epochs=0, rank=0) == 3
    log_full = dict(m_full.get_current_log())
    assert os.path.exists(tmp_path / "models" / "2_G.pth") and os.path.exists(tmp_path / "training_state" / "2.state")
    state = torch.load(tmp_path / "training_state" / "2.state", weights_only=False)
    assert state["iter"] == 2 and state["epoch"] == 0 and len(state["optimizers"][0]["state"]) == len(m_full.trainer.params) == 354
    # the optimizer entry is a valid torch.optim.Adam state_dict for the same parameter list
    ref_params = [torch.nn.Parameter(p.detach().cpu().clone()) for p in m_full.trainer.params]
    adam = torch.optim.Adam(ref_params, lr=1e-4)
    adam.load_state_dict(state["optimizers"][0])
    assert float(adam.state[ref_params[0]]["step"]) == 2.0
    # resume: fresh model from the step-2 weights + state, then the third step only
    assert state["noise"] == {"seed": 11, "offset": 2}                  # the eps stream position is part of the training state
    opt2, m_res = make(pretrain=str(tmp_path / "models" / "2_G.pth"))
    m_res.netG.module.set_noise(999, 0)                                  # a wrong stream: resume_training must restore (11, 2)
    last = train_loop.train(opt2, m_res, loader[2:], resume_state=state, total_epochs=0, rank=0)
    assert last == 3
    log_res = m_res.get_current_log()
    for k in ("l_forw_fit", "l_back_rec", "loss"):
        assert abs(log_res[k] - log_full[k]) <= 1e-5 * max(1.0, abs(log_full[k])), (k, log_res[k], log_full[k])
    a, b = m_full.netG.module.state_dict(), m_res.netG.module.state_dict()
    moved = max(float((a[k] - sd0[k]).abs().max()) for k in a)
    n_all = sum(v.numel() for v in a.values())
    n_off = sum(int(((a[k] - b[k]).abs() > 1e-6).sum()) for k in a)       # Adam steps of +-lr: only sign flips of noise-level
    worst = max(float((a[k] - b[k]).abs().max()) for k in a)              # gradients (float atomics in wgrad) can differ
    assert moved > 1e-5 and n_off <= 0.005 * n_all and worst <= 2.1e-4, (moved, n_off, n_all, worst)


# ------------------------------------------------------------------------------------------------ f3: 2x operators
def test_2x_operators_bit_exact(dev, golden_dir):
    """FrequencyAnalyzer(k=2) and HaarDownsampling kernels against the reference's modules (fixture) and the oracle."""
    from selfc_b200 import engine
    from selfc_b200.arch import FrequencyAnalyzer, HaarDownsampling
    g = np.load(os.path.join(golden_dir, "ops2x.npz"))
    fa2 = FrequencyAnalyzer(3, k=2)
    assert np.array_equal(fa2(_t(g["x"]).to(dev)).cpu().numpy(), g["fa2_fwd"])
    assert np.array_equal(fa2(_t(g["z15"]).to(dev), rev=True).cpu().numpy(), g["fa2_rev"])
    h3, h5 = HaarDownsampling(3).to(dev), HaarDownsampling(5).to(dev)
    assert sorted(h3.state_dict().keys()) == [str(k) for k in g["haar_state_keys"]]
    assert np.array_equal(h3(_t(g["x"]).to(dev)).cpu().numpy(), g["haar3_fwd"])
    assert np.array_equal(h5(_t(g["xh"]).to(dev)).cpu().numpy(), g["haar5_fwd"])
    assert np.array_equal(h3(_t(g["zh"]).to(dev), rev=True).cpu().numpy(), g["haar3_rev"])
    gen = torch.Generator().manual_seed(41)
    x = torch.rand(3, 3, 270, 482, generator=gen)
    assert torch.equal(engine.fa_forward(x.to(dev), 2).cpu(), so.fa_forward(x, 2))
    assert torch.equal(engine.haar_forward(x.to(dev)).cpu(), so.haar_forward(x))
    z = torch.randn(2, 15, 33, 47, generator=gen)
    assert torch.equal(engine.fa_reverse(z.to(dev), 2).cpu(), so.fa_reverse(z, 2))
    zh = torch.randn(2, 28, 33, 47, generator=gen)
    assert torch.equal(engine.haar_reverse(zh.to(dev)).cpu(), so.haar_reverse(zh))
    assert engine.haar_forward(torch.zeros(0, 3, 4, 4, device=dev)).shape == (0, 12, 2, 2)
    with pytest.raises(ValueError):
        engine.haar_forward(torch.zeros(1, 3, 5, 4, device=dev))
    with pytest.raises(ValueError):
        engine.fa_forward(torch.zeros(1, 3, 6, 6, device=dev), 3)


# ------------------------------------------------------------------------------------------------ a13 building blocks (training step)
# BF16X3 training: the forward (and the recomputed forward) runs on the tensor cores with 16-bit-mantissa operands, so an activation
# that the fp32 reference computes within ~1e-5 of zero can come out on the other side of zero and its LeakyReLU derivative flips
# (1 <-> 0.2) -- the sensitivity any re-ordered fp32 implementation has (cuDNN vs CPU), ten times likelier here.  A flip changes the
# gradient of ONE activation by 80 % and everything that back-propagates from it; on these few-hundred-pixel test clips that is a
# visible, LOCAL difference (measured: 0-4 flips per coupling block).  The bf16x3 comparisons are therefore held to a relative L2
# bound per tensor (and tight bounds on the losses and the gradient norms) instead of an element-wise maximum.
X3_GRAD_REL_L2 = 3e-2


def _assert_grad(got, ref, mode, atol, name="", floor=0.0):
    if mode == "bf16x3":
        err = float((got.double() - ref.double()).norm())
        bound = X3_GRAD_REL_L2 * float(ref.double().norm()) + floor * ref.numel() ** 0.5 + 1e-6
        assert err <= bound, f"{name}: relative L2 error {err / max(float(ref.double().norm()), 1e-30):.3e} (bound {X3_GRAD_REL_L2})"
    else:
        torch.testing.assert_close(got, ref, rtol=0, atol=atol, msg=lambda m, n=name: f"{n}: {m}")


@pytest.mark.parametrize("mode", ["fp32", "bf16x3"])
@pytest.mark.parametrize("prefix,cin,cout", [("operations.2.F", 48, 3), ("operations.6.G", 3, 48), ("stp_net.local_m2", 64, 64)])
def test_d2dt_backward_vs_autograd(dev, prefix, cin, cout, mode):
    """Backward of one dense block (dgrad = the forward implicit GEMM on flipped weights, wgrad = pixel reduction) against
    torch autograd on the oracle's D2DTInput: input gradient and all ten parameter gradients (both training modes)."""
    sd = so.make_state_dict(8)
    eng = _engine(dev, sd, mode)
    b, t, h, w = 2, 3, 11, 14
    gen = torch.Generator().manual_seed(31)
    x = torch.randn(b * t, cin, h, w, generator=gen) * 0.5
    gy = torch.randn(b * t, cout, h, w, generator=gen)
    leaf = {k: v.clone().requires_grad_(True) for k, v in sd.items() if k.startswith(prefix + ".")}
    xr = x.clone().requires_grad_(True)
    y = so.d2dt(leaf, prefix, xr, t)
    y.backward(gy)
    gx, grads = eng.d2dt_backward(prefix, x.to(dev), gy.to(dev), t)
    _assert_grad(gx.cpu(), xr.grad, mode, 1e-4 + 1e-4 * float(xr.grad.abs().max()), "gx")
    for name, gval in grads.items():
        ref = leaf[name].grad
        tol = 2e-4 * float(ref.abs().max()) + 1e-5
        _assert_grad(gval.cpu(), ref, mode, tol, name)


@pytest.mark.parametrize("mode", ["fp32", "bf16x3"])
@pytest.mark.parametrize("rev", [False, True])
def test_invblock_backward_vs_autograd(dev, rev, mode):
    """Backward of one affine-coupling block in both directions (recompute + three dense-block backwards + the coupling
    arithmetic) against autograd on the oracle's InvBlockExp."""
    sd = so.make_state_dict(3, gain=1.5)
    eng = _engine(dev, sd, mode)
    b, t, h, w, blk = 2, 3, 9, 12, 4
    prefix = f"operations.{blk + 1}"
    gen = torch.Generator().manual_seed(5)
    z = torch.randn(b * t, 51, h, w, generator=gen) * 0.5
    gz = torch.randn(b * t, 51, h, w, generator=gen)
    leaf = {k: v.clone().requires_grad_(True) for k, v in sd.items() if k.startswith(prefix + ".")}
    zr = z.clone().requires_grad_(True)
    out = (so.invblock_reverse if rev else so.invblock_forward)(leaf, prefix, zr, t)
    out.backward(gz)
    gzin, grads = eng.invblock_backward(blk, rev, z.to(dev), gz.to(dev), t)
    _assert_grad(gzin.cpu(), zr.grad, mode, 2e-4 + 2e-4 * float(zr.grad.abs().max()), "gz")
    for name, gval in grads.items():
        ref = leaf[name].grad
        tol = 3e-4 * float(ref.abs().max()) + 1e-5
        _assert_grad(gval.cpu(), ref, mode, tol, name)


@pytest.mark.parametrize("mode", ["fp32", "bf16x3"])
def test_head_sampler_backward_vs_autograd(dev, mode):
    """Backward of tail_gmm (three 1x1 convs + LeakyReLUs) and the soft-GMM sampler (softmax over hf, clamp, exp, injected
    eps) against autograd on the oracle.  bf16x3: the recomputed forward and the input gradients as tensor-core pointwise convs, the
    weight gradients through the pointwise form of the tensor-core weight-gradient kernel (relative-L2 bound: LeakyReLU mask flips)."""
    sd = so.make_state_dict(12, gain=2.0)
    eng = _engine(dev, sd, mode)
    b, t, h, w = 2, 3, 7, 10
    gen = torch.Generator().manual_seed(9)
    feat = torch.randn(b * t, 64, h, w, generator=gen)
    gv = torch.randn(b * t, 48, h, w, generator=gen)
    eps = so.make_eps(b, t, h, w, 33)
    leaf = {k: v.clone().requires_grad_(True) for k, v in sd.items() if k.startswith("stp_net.tail_gmm.")}
    fr = feat.clone().requires_grad_(True)
    v = so.gmm_sample(so.gmm_head(leaf, fr), eps, t)
    v.backward(gv)
    gfeat, grads = eng.head_sampler_backward(feat.to(dev), gv.to(dev), t, eps=eps.to(dev))
    _assert_grad(gfeat.cpu(), fr.grad, mode, 3e-4 * float(fr.grad.abs().max()) + 1e-6, "gfeat")
    for name, gval in grads.items():
        ref = leaf[name].grad
        _assert_grad(gval.cpu(), ref, mode, 3e-4 * float(ref.abs().max()) + 1e-6, name)


@pytest.mark.parametrize("mode", ["fp32", "bf16x3"])
@pytest.mark.parametrize("h,w,t", [(10, 18, 2), (33, 40, 7)])
def test_global_agg_backward_vs_autograd(dev, h, w, t, mode):
    """Backward of GlobalAgg (pooled descriptor -> T x T softmax mixing -> proj1 mix + residual) against autograd on the
    oracle: input gradient and the eight parameter gradients (fc through the overlapping adaptive-pool bins); bf16x3: proj1's
    weight gradient on the tensor cores."""
    sd = so.make_state_dict(6, gain=2.0)
    eng = _engine(dev, sd, mode)
    b = 2
    prefix = "stp_net.global_m2"
    gen = torch.Generator().manual_seed(h)
    x = torch.randn(b * t, 64, h, w, generator=gen)
    gout = torch.randn(b * t, 64, h, w, generator=gen)
    leaf = {k: v.clone().requires_grad_(True) for k, v in sd.items() if k.startswith(prefix + ".")}
    xr = x.clone().requires_grad_(True)
    so.global_agg(leaf, prefix, xr, t).backward(gout)
    gx, grads = eng.global_agg_backward(prefix, x.to(dev), gout.to(dev), t)
    torch.testing.assert_close(gx.cpu(), xr.grad, rtol=0, atol=3e-4 * float(xr.grad.abs().max()) + 1e-6)
    for name, gval in grads.items():
        ref = leaf[name].grad
        # proj3.bias shifts every logit of a softmax row equally: its exact gradient is 0 and both sides return fp32 noise (~1e-6)
        torch.testing.assert_close(gval.cpu(), ref, rtol=0, atol=5e-4 * float(ref.abs().max()) + 2e-5)


@pytest.mark.parametrize("mode", ["fp32", "bf16x3"])
def test_train_step_gradients_vs_reference_fixture_and_oracle(dev, golden_dir, mode):
    """Row a13: losses and all 354 parameter gradients of one training step (down, STE quantiser, STP + sampler, up, two
    losses, x 144*144*3) against the gradients the REFERENCE's own modules produced (tests/golden/train_t3.npz) and against
    autograd on the oracle.  bf16x3: the same step with the forward, the recomputed forward and the input-gradient / weight-gradient
    convolutions on the tcgen05 kernels ((hi, lo) operands) -- held to the SAME tolerances."""
    g = np.load(os.path.join(golden_dir, "train_t3.npz"))
    b, t, hh, ww, wseed, xseed = [int(v) for v in g["meta"]]
    sd = so.make_state_dict(wseed)
    eng = _engine(dev, sd, mode)
    x = so.make_frames(b, t, hh, ww, xseed)
    eps = so.make_eps(b, t, hh // 4, ww // 4, int(g["eps_seed"]))
    ref_l = _t(g["ref_l"])
    grads, losses = eng.train_grads(x.to(dev), ref_l.to(dev), t, eps=eps.to(dev))
    losses = losses.cpu()
    assert abs(losses[1].item() - float(g["l_forw"])) <= 2e-5 * abs(float(g["l_forw"])) + 1e-7
    assert abs(losses[2].item() - float(g["l_back"])) <= 2e-5 * abs(float(g["l_back"])) + 1e-7
    assert abs(losses[0].item() - float(g["loss"])) <= 5e-5 * abs(float(g["loss"]))
    names = [str(n) for n in g["names"]]
    norms = np.array([float(grads[n].double().norm()) for n in names])
    np.testing.assert_allclose(norms, g["grad_norms"], rtol=5e-3, atol=1e-2)
    ogr, _, _, _ = so.train_grads(sd, x, ref_l, eps, t)
    gmax = max(float(v.abs().max()) for v in ogr.values())
    for n in names:
        ref = ogr[n]
        # fp32 sums over all pixels in a different order than CPU autograd, through 16 coupling blocks: 0.5 % of the tensor's
        # largest gradient (the gradient norms above are held to 0.5 % against the reference's own numbers)
        tol = 5e-3 * float(ref.abs().max()) + 1e-5 * gmax
        _assert_grad(grads[n].cpu(), ref, mode, tol, n, floor=1e-5 * gmax)


def _make_net(dev, sd):
    from selfc_b200 import networks, options
    here = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    opt = options.dict_to_nonedict(options.parse(os.path.join(here, "selfc_b200", "configs", "selfc_large_synthetic.yml"), is_train=False))
    net = networks.define_G(opt)
    net.load_state_dict(sd, strict=True)
    return net.to(dev)


def test_optimizer_step_matches_torch_adam_with_clipping(dev, golden_dir):
    """clip_grad_norm_(10) + Adam(1e-4, wd 1e-14) through selfc_adam_step on the flat gradient against torch.optim.Adam on the
    oracle's autograd gradients: every parameter after one step, and after a second step on the updated weights."""
    from selfc_b200.train import Trainer
    from selfc_b200.global_var import GlobalVar
    g = np.load(os.path.join(golden_dir, "train_t3.npz"))
    b, t, hh, ww, wseed, xseed = [int(v) for v in g["meta"]]
    sd = so.make_state_dict(wseed)
    x = so.make_frames(b, t, hh, ww, xseed)
    eps = so.make_eps(b, t, hh // 4, ww // 4, int(g["eps_seed"]))
    ref_l = _t(g["ref_l"])
    GlobalVar.set_Temporal_LEN(t)
    net = _make_net(dev, sd)
    tr = Trainer(net, dev, lr=1e-4, betas=(0.9, 0.999), weight_decay=1e-14, max_norm=10.0)
    # CPU reference: the same two steps with torch
    ref_p = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    opt = torch.optim.Adam(list(ref_p.values()), lr=1e-4, betas=(0.9, 0.999), weight_decay=1e-14)
    for it in range(2):
        losses = tr.step(x.to(dev), ref_l.to(dev), t, eps=eps.to(dev))
        opt.zero_grad()
        loss, _, _ = so.train_losses(ref_p, x, ref_l, eps, t)
        loss.backward()
        torch.nn.utils.clip_grad_norm_(list(ref_p.values()), 10.0)
        opt.step()
        assert abs(losses[0].item() - loss.item()) <= 2e-4 * abs(loss.item())
        own = dict(net.named_parameters())
        n_bad = n_all = 0
        for k, v in ref_p.items():
            # Adam's first steps move every weight by ~lr * sign(g): compare the UPDATE to 2 % of lr; where the gradient is at
            # fp32-noise level its sign (hence an update of +-lr) is arbitrary on both sides -- allow 0.2 % such elements over the whole model
            diff = (own[k].detach().cpu() - v.detach()).abs()
            assert diff.max().item() <= (it + 1) * 2.1e-4, f"{k} step {it}: max |diff| {diff.max().item():.3e}"
            n_bad += int((diff > (it + 1) * 2e-6).sum().item())
            n_all += diff.numel()
        assert n_bad <= 2e-3 * n_all, f"step {it}: {n_bad} of {n_all} elements differ by more than 2 % of lr"


def test_model_optimize_parameters_matches_oracle_step(dev):
    """SelfCModel (training options) -> feed_data -> optimize_parameters: the logged losses are the oracle's, the weights
    move as torch.optim.Adam would move them, update_learning_rate follows the MultiStepLR milestones."""
    from selfc_b200 import model as smodel, options
    from selfc_b200.global_var import GlobalVar
    here = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    opt = options.dict_to_nonedict(options.parse(os.path.join(here, "selfc_b200", "configs", "selfc_large_train_synthetic.yml"),
                                                 is_train=True))
    torch.manual_seed(10)
    m = smodel.create_model(opt)
    net = m.netG.module
    sd = {k: v.detach().cpu().clone() for k, v in net.state_dict().items()}
    b, t, hh, ww = 1, 3, 32, 48
    GlobalVar.set_Temporal_LEN(t)
    x = so.make_frames(b, t, hh, ww, 77)
    eps = so.make_eps(b, t, hh // 4, ww // 4, 3)
    net.inject_eps(eps.to(dev))
    m.feed_data({"GT": x.reshape(b, t, 3, hh, ww).transpose(1, 2)})
    m.optimize_parameters(1)
    ref_p = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    loss, l_forw, l_back = so.train_losses(ref_p, x, so.gaussian_downsample(x), eps, t)
    log = m.get_current_log()
    assert abs(log["loss"] - loss.item()) <= 2e-4 * abs(loss.item())
    assert abs(log["l_forw_fit"] - l_forw.item()) <= 1e-5 and abs(log["l_back_rec"] - l_back.item()) <= 1e-5
    moved = max(float((p.detach().cpu() - sd[k]).abs().max()) for k, p in net.named_parameters())
    assert 0.5e-4 <= moved <= 1.05e-4                           # Adam's first step: |update| <= lr
    # forward builds no autograd graph: in train mode with grad enabled it must say so instead of returning graph-less tensors
    with pytest.raises(RuntimeError, match="no autograd graph"):
        net(x=x.to(dev))
    # the next forward uses the updated weights (packed images were invalidated)
    with torch.no_grad():
        out, _ = net(x=x.to(dev))
        ref_out = so.net_down({k: p.detach().cpu() for k, p in net.named_parameters()}, x, t)
    torch.testing.assert_close(out.cpu(), ref_out, rtol=0, atol=2e-4)
    m.update_learning_rate(100000)
    assert m.get_current_learning_rate() == 5e-5
