// libselfc_b200: context (packed-weight cache), workspace layout, orchestration of the rescaling path and the
// C-ABI declared in include/selfc_b200.h.
//
// Orchestration restates SelfCInvNet.forward (models/modules/SelfC_GMM_arch_inv.py:450-490), InvBlockExp.forward
// (:21-33), D2DTInput.forward (Subnet_constructor.py:115-133) and STPNet.forward (:358-394) as a fixed launch
// sequence over pixel-major buffers; nothing here is a translation of reference code.
#include <memory>
#include <mutex>
#include <new>
#include <type_traits>
#include <string.h>
#include <vector>

#include "common.cuh"
#include "kernels.h"
#include "conv_tc.h"
#include "tc_ptx.cuh"
#include "net_ctx.h"
#include "wgrad_tc.h"

namespace selfc {

// ---- per-thread error text + launch counter --------------------------------------------------------------
static thread_local char g_err[512] = "";
static thread_local uint64_t g_launches = 0;
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
uint64_t& launch_counter() { return g_launches; }


}  // namespace selfc

using namespace selfc;

namespace selfc {

// ---- per-launch profiler ----------------------------------------------------------------------------------
static void prof_begin(const selfc_ctx* cctx, cudaStream_t st, int cls, double work) {
  selfc_ctx* ctx = const_cast<selfc_ctx*>(cctx);
  if (!ctx->prof_on) return;
  ProfRec r;
  r.cls = cls;
  r.work = work;
  cudaEventCreate(&r.a);
  cudaEventCreate(&r.b);
  cudaEventRecord(r.a, st);
  ctx->prof.push_back(r);
}
static void prof_end(const selfc_ctx* cctx, cudaStream_t st) {
  selfc_ctx* ctx = const_cast<selfc_ctx*>(cctx);
  if (!ctx->prof_on || ctx->prof.empty()) return;
  cudaEventRecord(ctx->prof.back().b, st);
}
#define PROF(ctx, st, cls, work, call) \
  do {                                 \
    prof_begin(ctx, st, cls, work);    \
    int rc_p = (call);                 \
    prof_end(ctx, st);                 \
    if (rc_p != 0) return rc_p;        \
  } while (0)

// element types of the tensor-core modes: bf16 (BF16 mode) and (hi, lo) bf16 pairs (BF16X3 mode, common.cuh)
template <typename T> constexpr bool kTc = std::is_same<T, __nv_bfloat16>::value || std::is_same<T, bfx2>::value;
template <typename T> constexpr bool kX2 = std::is_same<T, bfx2>::value;
template <typename T> static __nv_bfloat16* as_bf(T* p) { return reinterpret_cast<__nv_bfloat16*>(p); }
template <typename T> static const __nv_bfloat16* as_bf(const T* p) { return reinterpret_cast<const __nv_bfloat16*>(p); }
static bool mode_tc(const selfc_ctx* ctx) { return ctx->mode == SELFC_MODE_BF16 || ctx->mode == SELFC_MODE_BF16X3; }

Workspace make_workspace(const selfc_ctx* ctx, int B, int T, int h, int w) {
  Workspace ws;
  const size_t M = (size_t)B * T * h * w;
  const size_t es = ctx->mode == SELFC_MODE_BF16 ? 2 : 4;
  // slack after every region: conv3x3's position-pair tensor map reads one position (32 bytes) before and after a dense buffer
  // (never used in an output); `z` comes first, so no dense buffer starts the workspace
  const size_t guard = 4096;
  size_t off = 0;
  auto take = [&](size_t bytes) {
    size_t o = off;
    off = align_up(off + bytes + guard, 1024);
    return o;
  };
  ws.gpitch = ctx->xpad3 + 4 * kGrowth;
  ws.z = take(M * kZQuads * 16);
  ws.sbuf = take(M * kHF * 4);
  ws.fbuf = take(M * ws.fpitch * es);
  ws.gbuf = take(M * ws.gpitch * es);
  ws.hbuf = take(M * ws.gpitch * es);
  ws.stpbuf = take(M * ws.spitch * es);
  ws.feat = take(M * kStpC * es);
  ws.fact = take(M * kStpC * es);
  ws.h1 = take(M * 128 * es);
  ws.h2 = take(M * 256 * es);
  ws.params = take(M * 720 * 4);
  ws.nsplit = (int)((size_t)h * w / 512);     // ga_stat: enough CTAs (nsplit x B*T) to cover the DRAM latency
  if (ws.nsplit < 1) ws.nsplit = 1;
  if (ws.nsplit > 128) ws.nsplit = 128;
  ws.wmap = take((size_t)h * w * 4);
  ws.partial = take((size_t)B * T * ws.nsplit * 64 * 4);
  ws.wmat = take((size_t)B * T * T * 4);
  ws.wsum = take((size_t)B * T * 4);
  ws.lrq = take(M * 3 * 4);                   // 8-bit entry points: the LR frames as fp32 NCHW between the two halves
  ws.total = off;
  return ws;
}

// ---- dense block: conv1..4 in place, then conv5 with the given epilogue ---------------------------------------
template <typename T>
static int run_dense_convs(const selfc_ctx* ctx, const DenseW& W, T* buf, int pitch, const Dims& d, cudaStream_t st, int k_first = 0,
                           int k_last = 3, const DenseW* W2 = nullptr, T* buf2 = nullptr) {
  const long long slabM = dense_slab(ctx, d);
  if constexpr (std::is_same<T, __nv_bfloat16>::value) {
    // the whole block's conv1..conv4 (conv1..conv3 for cin = 64) as ONE launch with the growth channels kept on chip
    // (dense_fused.cu); SELFC_DB_FUSED=0 selects the layer-by-layer launches (bit-identical results, A/B and parity tests)
    static int fused_on = -1;
    if (fused_on < 0) {
      const char* e = getenv("SELFC_DB_FUSED");
      fused_on = (e && atoi(e) == 0) ? 0 : 1;
    }
    const bool dual = W2 != nullptr && W2->tc[0].img_pair != nullptr;
    if (fused_on && ctx->mode == SELFC_MODE_BF16 && k_first == 0 && k_last == 3 && W.tc[0].img_pair != nullptr && (W2 == nullptr || dual)) {
      const int L = dense_fused_layers(W.xpad);
      if (L > 0) {
        double flops = 0.0;
        for (int k = 0; k < L; ++k) flops += 2.0 * (double)d.M() * 9.0 * (W.cin + kGrowth * k) * kGrowth;
        PROF(ctx, st, 0, (dual ? 2.0 : 1.0) * flops,
             launch_dense_fused(W.tc, L, reinterpret_cast<__nv_bfloat16*>(buf), slabM, W.xpad, d.B * d.T, d.h, d.w, st, dual ? W2->tc : nullptr,
                                dual ? reinterpret_cast<__nv_bfloat16*>(buf2) : nullptr));
        k_first = L;
      }
    }
  }
  for (int k = k_first; k <= k_last; ++k) {
    const int cin = W.xpad + kGrowth * k;
    const double flops = 2.0 * (double)d.M() * 9.0 * (W.cin + kGrowth * k) * kGrowth;   // algorithmic (unpadded) FLOPs
    if (kTc<T> && mode_tc(ctx) && W.tc[k].img != nullptr) {
      if (W2 != nullptr && W2->tc[k].img != nullptr) {
        // two dense blocks of the same shape (G and H of a coupling) side by side: one launch per layer for both
        PROF(ctx, st, 0, 2.0 * flops, launch_conv3x3_tc(W.tc[k], reinterpret_cast<__nv_bfloat16*>(buf), slabM, cin, /*out_off=*/cin, d.B * d.T,
                                                        d.h, d.w, st, &W2->tc[k], reinterpret_cast<__nv_bfloat16*>(buf2), kX2<T>));
        continue;
      }
      PROF(ctx, st, 0, flops, launch_conv3x3_tc(W.tc[k], reinterpret_cast<__nv_bfloat16*>(buf), slabM, cin, /*out_off=*/cin, d.B * d.T, d.h, d.w, st,
                                                nullptr, nullptr, kX2<T>));
      continue;
    }
    ConvArgs<T> a;
    a.in = buf; a.in_pitch = pitch; a.cin = cin; a.in_slabM = slabM;
    a.w = W.w[k]; a.bias = W.b[k]; a.np = W.np[k]; a.cout = kGrowth;
    a.taps = 9; a.tap_mode = TAP_SPATIAL;
    a.BT = d.B * d.T; a.Tn = d.T; a.h = d.h; a.w_ = d.w;
    a.epi = EPI_STORE; a.act = 1;
    a.outT = buf; a.outT_pitch = pitch; a.outT_off = cin; a.outT_slabM = slabM;
    PROF(ctx, st, 0, flops, launch_conv_simt<T>(a, st));
  }
  return 0;
}

template <typename T>
static ConvArgs<T> conv5_args(const selfc_ctx* ctx, const DenseW& W, const T* buf, int pitch, const Dims& d) {
  ConvArgs<T> a;
  a.in = buf; a.in_pitch = pitch; a.cin = W.xpad + 4 * kGrowth; a.in_slabM = dense_slab(ctx, d);
  a.w = W.w[4]; a.bias = W.b[4]; a.np = W.np[4]; a.cout = W.cout;
  a.taps = 3; a.tap_mode = TAP_TEMPORAL;
  a.BT = d.B * d.T; a.Tn = d.T; a.h = d.h; a.w_ = d.w;
  return a;
}

static double conv5_flops(const DenseW& W, const Dims& d) { return 2.0 * (double)d.M() * 3.0 * (W.cin + 4 * kGrowth) * W.cout; }

// conv5 / GlobalAgg-apply dispatch: tcgen05 kernel in BF16 mode when T accumulators fit TMEM, fp32-FMA kernel otherwise
template <typename T>
static int launch_temporal(const selfc_ctx* ctx, const TcTempW& tw, const ConvArgs<T>& a, const Dims& d, cudaStream_t st) {
  if constexpr (kTc<T>) {
    if (mode_tc(ctx) && temporal_tc_supported(tw, d.T)) {
      TcTempArgs t;
      t.in = as_bf(a.in); t.in_pitch = a.in_pitch; t.B = d.B; t.T = d.T; t.hw = (int)d.hw(); t.in_slabM = a.in_slabM;
      t.epi = a.epi; t.rev = a.rev; t.act = a.act;
      t.outT = as_bf(a.outT); t.outT_pitch = a.outT_pitch; t.outT_off = a.outT_off; t.outT_slabM = a.outT_slabM; t.copy_slabM = a.copy_slabM;
      t.outF = a.outF; t.outF_pitch = a.outF_pitch;
      t.z = a.z; t.sbuf = a.sbuf;
      t.copyA = as_bf(a.copyA); t.copyA_pitch = a.copyA_pitch; t.copyB = as_bf(a.copyB); t.copyB_pitch = a.copyB_pitch; t.copy_pad = a.copy_pad;
      t.wmat = a.wmat; t.wsum = a.wsum; t.resid = as_bf(a.resid); t.resid_pitch = a.resid_pitch;
      t.outAct = as_bf(a.outAct); t.outAct_pitch = a.outAct_pitch;
      const int rc = launch_temporal_tc(tw, t, st);
      if (rc != SELFC_E_UNSUPPORTED) return rc;      // shared memory cannot hold the frame ring: fp32-FMA kernel below
    }
  }
  return launch_conv_simt<T>(a, st);
}

// InvBlockExp (SelfC_GMM_arch_inv.py:21-33) on the latent state z, forward or reverse
template <typename T>
static int run_invblock(const selfc_ctx* ctx, int blk, bool rev, char* wsp, const Workspace& ws, const Dims& d, cudaStream_t st,
                        const BlockBufsV* bb = nullptr) {
  float* z = reinterpret_cast<float*>(wsp + ws.z);
  float* sbuf = bb ? bb->s : reinterpret_cast<float*>(wsp + ws.sbuf);
  T* fbuf = bb ? static_cast<T*>(bb->f) : reinterpret_cast<T*>(wsp + ws.fbuf);
  T* gbuf = bb ? static_cast<T*>(bb->g) : reinterpret_cast<T*>(wsp + ws.gbuf);
  T* hbuf = bb ? static_cast<T*>(bb->h) : reinterpret_cast<T*>(wsp + ws.hbuf);
  // where the block's last epilogue leaves the next block's input (training keeps every block's buffers; see BlockBufsV)
  T* f_out = bb && !rev ? static_cast<T*>(bb->f_next) : fbuf;      // Y2's copy: forward -> the next block's F; reverse -> this block's F
  T* g_out = bb && rev ? static_cast<T*>(bb->g_next) : gbuf;       // Y1's copies: forward -> this block's G, H; reverse -> the next block's
  T* h_out = bb && rev ? static_cast<T*>(bb->h_next) : hbuf;
  const DenseW& F = ctx->inv[blk][0];
  const DenseW& G = ctx->inv[blk][1];
  const DenseW& H = ctx->inv[blk][2];
  auto do_F = [&]() -> int {
    if constexpr (std::is_same<T, __nv_bfloat16>::value) {
      // F has 3 outputs: conv5's temporal taps are applied inside the fused conv1..4 launch while the block's activations are on
      // chip (9 partial products per pixel), x1..x4 never go to HBM, and a small kernel finishes conv5 + the additive coupling.
      // SELFC_F5=0: conv1..4 launch + the temporal kernel (what G / H / the STP blocks use).
      static int f5_on = -1;
      if (f5_on < 0) {
        const char* e = getenv("SELFC_F5");
        const char* e2 = getenv("SELFC_DB_FUSED");
        f5_on = ((e && atoi(e) == 0) || (e2 && atoi(e2) == 0)) ? 0 : 1;
      }
      if (f5_on && ctx->mode == SELFC_MODE_BF16 && F.f5img != nullptr && F.tc[0].img_pair != nullptr && F.xpad == 48 &&
          dense_fused_layers(F.xpad) == 4) {
        double flops = conv5_flops(F, d);
        for (int k = 0; k < 4; ++k) flops += 2.0 * (double)d.M() * 9.0 * (F.cin + kGrowth * k) * kGrowth;
        PROF(ctx, st, 0, flops,
             launch_dense_fused(F.tc, 4, reinterpret_cast<__nv_bfloat16*>(fbuf), dense_slab(ctx, d), F.xpad, d.B * d.T, d.h, d.w, st, nullptr, nullptr,
                                F.f5img, sbuf));
        PROF(ctx, st, 1, 0.0,
             launch_f5_combine(sbuf, F.t5.bias, z, reinterpret_cast<__nv_bfloat16*>(g_out), reinterpret_cast<__nv_bfloat16*>(h_out), d.T, d.hw(), d.M(),
                               rev ? 1 : 0, st));
        return 0;
      }
    }
    SELFC_TRY(run_dense_convs<T>(ctx, F, fbuf, ws.fpitch, d, st));
    ConvArgs<T> a = conv5_args<T>(ctx, F, fbuf, ws.fpitch, d);
    a.epi = EPI_COUPLE_Y1; a.rev = rev ? 1 : 0; a.z = z;
    a.copyA = g_out; a.copyA_pitch = ws.gpitch; a.copyB = h_out; a.copyB_pitch = ws.gpitch; a.copy_pad = ctx->xpad3;
    a.copy_slabM = dense_slab(ctx, d);
    PROF(ctx, st, 1, conv5_flops(F, d), launch_temporal<T>(ctx, F.t5, a, d, st));
    return 0;
  };
  auto do_HG = [&]() -> int {
    // H and G read the same X slab (y1 / x1) and have the same shape: in BF16 mode their conv1..4 share launches
    static int dual_on = -1;     // SELFC_DUAL_GH=0: separate launches for H and G (A/B)
    if (dual_on < 0) {
      const char* e = getenv("SELFC_DUAL_GH");
      dual_on = (e && atoi(e) == 0) ? 0 : 1;
    }
    const bool dual = dual_on && kTc<T> && mode_tc(ctx) && H.tc[0].img != nullptr && G.tc[0].img != nullptr;
    if (dual) SELFC_TRY(run_dense_convs<T>(ctx, H, hbuf, ws.gpitch, d, st, 0, 3, &G, gbuf));
    else SELFC_TRY(run_dense_convs<T>(ctx, H, hbuf, ws.gpitch, d, st));
    if constexpr (std::is_same<T, __nv_bfloat16>::value) {
      static int fuse_hg = -1;     // SELFC_FUSE_HG=0: separate conv5 launches for H (-> s) and G (A/B)
      if (fuse_hg < 0) {
        const char* e = getenv("SELFC_FUSE_HG");
        fuse_hg = (e && atoi(e) == 0) ? 0 : 1;
      }
      if (dual && fuse_hg && temporal_tc_supported(H.t5, d.T) && temporal_tc_supported(G.t5, d.T)) {
        // conv5 of H and of G in ONE launch: s = 2*sigmoid(H(y1))-1 is consumed from the accumulator, never written to HBM
        TcTempArgs t;
        t.in = hbuf; t.in2 = gbuf; t.in_pitch = ws.gpitch; t.B = d.B; t.T = d.T; t.hw = (int)d.hw(); t.in_slabM = dense_slab(ctx, d);
        t.epi = EPI_COUPLE_HG; t.rev = rev ? 1 : 0; t.z = z;
        t.copyA = f_out; t.copyA_pitch = ws.fpitch; t.copy_slabM = dense_slab(ctx, d);
        prof_begin(ctx, st, 1, conv5_flops(H, d) + conv5_flops(G, d));
        const int rc = launch_temporal_tc(H.t5, t, st, &G.t5);
        prof_end(ctx, st);
        if (rc == 0) return 0;
        if (rc != SELFC_E_UNSUPPORTED) return rc;
      }
    }
    ConvArgs<T> a = conv5_args<T>(ctx, H, hbuf, ws.gpitch, d);
    a.epi = EPI_COUPLE_S; a.sbuf = sbuf;
    PROF(ctx, st, 1, conv5_flops(H, d), launch_temporal<T>(ctx, H.t5, a, d, st));
    if (!dual) SELFC_TRY(run_dense_convs<T>(ctx, G, gbuf, ws.gpitch, d, st));
    ConvArgs<T> g = conv5_args<T>(ctx, G, gbuf, ws.gpitch, d);
    g.epi = EPI_COUPLE_Y2; g.rev = rev ? 1 : 0; g.z = z; g.sbuf = sbuf;
    g.copyA = f_out; g.copyA_pitch = ws.fpitch; g.copy_slabM = dense_slab(ctx, d);
    PROF(ctx, st, 1, conv5_flops(G, d), launch_temporal<T>(ctx, G.t5, g, d, st));
    return 0;
  };
  if (!rev) {
    SELFC_TRY(do_F());
    SELFC_TRY(do_HG());
  } else {
    SELFC_TRY(do_HG());
    SELFC_TRY(do_F());
  }
  return 0;
}

// GlobalAgg (SelfC_GMM_arch_inv.py:265-285): x = feat [M][64]; result -> outT (pitch/off) and/or outF
template <typename T>
static int run_global_agg(const selfc_ctx* ctx, const GaW& g, const T* feat, T* outT, int outT_pitch, long long outT_slabM, float* outF,
                          int outF_pitch, float* wmat_copy, char* wsp, const Workspace& ws, const Dims& d, cudaStream_t st, T* outAct = nullptr) {
  float* wmap = reinterpret_cast<float*>(wsp + ws.wmap);
  float* partial = reinterpret_cast<float*>(wsp + ws.partial);
  float* wmat = reinterpret_cast<float*>(wsp + ws.wmat);
  float* wsum = reinterpret_cast<float*>(wsp + ws.wsum);
  const double px_bytes = (double)d.M() * kStpC * sizeof(T);
  {
    // the pooling weight map: cached per (weights, h, w, stream) in the context instead of one launch per call
    GaW& gm = const_cast<GaW&>(g);
    const size_t need = (size_t)d.h * d.w * sizeof(float);
    if (gm.wmap_cache == nullptr || gm.wmap_cap < need) {
      if (gm.wmap_cache) cudaFree(gm.wmap_cache);
      gm.wmap_cache = nullptr;
      gm.wmap_cap = 0;
      gm.wmap_h = gm.wmap_w = 0;
      SELFC_CUDA(cudaMalloc(&gm.wmap_cache, need));
      gm.wmap_cap = need;
    }
    if (gm.wmap_h != d.h || gm.wmap_w != d.w || gm.wmap_stream != st) {
      PROF(ctx, st, 2, 0.0, launch_ga_wmap(g.fcw, gm.wmap_cache, d.h, d.w, st));
      gm.wmap_h = d.h;
      gm.wmap_w = d.w;
      gm.wmap_stream = st;
    }
    wmap = gm.wmap_cache;
  }
  PROF(ctx, st, 2, px_bytes, launch_ga_stat<T>(feat, kStpC, wmap, partial, ws.nsplit, d.B * d.T, (int)d.hw(), st));
  PROF(ctx, st, 2, 0.0, launch_ga_weights(partial, ws.nsplit, g.fcb, g.p2w, g.p2b, g.p3w, g.p3b, wmat, wsum, d.B, d.T, st));
  if (wmat_copy) SELFC_CUDA(cudaMemcpyAsync(wmat_copy, wmat, (size_t)d.B * d.T * d.T * 4, cudaMemcpyDeviceToDevice, st));
  if constexpr (kTc<T>) {
    if (mode_tc(ctx) && g.tp.img != nullptr) {
      // P = proj1(x) + bias as a tcgen05 pointwise GEMM over all M pixels (two pseudo-frames), then the T x T mix at full occupancy
      __nv_bfloat16* P = reinterpret_cast<__nv_bfloat16*>(wsp + ws.h1);
      TcTempArgs t;
      t.in = as_bf(feat); t.in_pitch = kStpC; t.B = 1; t.T = 2; t.hw = (int)((d.M() + 1) / 2); t.m_limit = d.M();
      t.epi = EPI_STORE; t.act = 0; t.outT = P; t.outT_pitch = kStpC;
      PROF(ctx, st, 2, 2.0 * px_bytes, launch_temporal_tc(g.tp, t, st));
      PROF(ctx, st, 2, 3.0 * px_bytes, launch_ga_mix(P, as_bf(feat), wmat, as_bf(outT), outT_pitch, outT_slabM, outF, outF_pitch, as_bf(outAct), d.B,
                                                     d.T, d.hw(), st, kX2<T>));
      return 0;
    }
  }
  ConvArgs<T> a;
  a.in = feat; a.in_pitch = kStpC; a.cin = kStpC;
  a.w = g.p1w; a.bias = g.p1b; a.np = 64; a.cout = 64;
  a.taps = 1; a.tap_mode = TAP_TMIX;
  a.BT = d.B * d.T; a.Tn = d.T; a.h = d.h; a.w_ = d.w;
  a.wmat = wmat; a.wsum = wsum;
  a.epi = EPI_GA; a.resid = feat; a.resid_pitch = kStpC;
  a.outT = outT; a.outT_pitch = outT_pitch; a.outT_off = 0; a.outT_slabM = outT_slabM;
  a.outF = outF; a.outF_pitch = outF_pitch; a.outF_off = 0;
  a.outAct = outAct; a.outAct_pitch = kStpC;
  PROF(ctx, st, 2, 2.0 * px_bytes, launch_temporal<T>(ctx, g.tp, a, d, st));
  return 0;
}

template <typename T>
static int down_impl(selfc_ctx* ctx, const float* hr, float* out51, uint8_t* lr_u8, float* lr_q, const Dims& d, char* wsp,
                     const Workspace& ws, cudaStream_t st, const uint8_t* hr_img = nullptr) {
  float* z = reinterpret_cast<float*>(wsp + ws.z);
  T* fbuf = reinterpret_cast<T*>(wsp + ws.fbuf);
  if (hr_img != nullptr)
    PROF(ctx, st, 5, (double)d.M() * (16 * 3 + 51 * 4), launch_fa_fwd_z_u8<T>(hr_img, z, fbuf, ws.fpitch, dense_slab(ctx, d), d.B * d.T, d.h, d.w, st));
  else
    PROF(ctx, st, 5, (double)d.M() * (16 * 3 * 4 + 51 * 4), launch_fa_fwd_z<T>(hr, z, fbuf, ws.fpitch, dense_slab(ctx, d), d.B * d.T, d.h, d.w, st));
  for (int blk = 0; blk < 8; ++blk) SELFC_TRY(run_invblock<T>(ctx, blk, false, wsp, ws, d, st));
  PROF(ctx, st, 5, (double)d.M() * (51 * 4 + (out51 ? 51 * 4 : 0) + 3), launch_export_down(z, out51, lr_u8, lr_q, d.M(), d.hw(), st));
  return 0;
}

template <typename T>
static int up_impl(selfc_ctx* ctx, const float* lr, const float* eps, uint64_t seed, uint64_t offset, float* hr, float* hf,
                   const Dims& d, char* wsp, const Workspace& ws, cudaStream_t st, const TrainHooks* hooks = nullptr,
                   uint8_t* hr_img = nullptr) {
  float* z = reinterpret_cast<float*>(wsp + ws.z);
  // the X slots of the first reverse block (block 8): the workspace buffers, or that block's own buffers when training keeps them
  const BlockBufsV* ub = hooks ? hooks->up_bufs : nullptr;
  T* gbuf = ub ? static_cast<T*>(ub[7].g) : reinterpret_cast<T*>(wsp + ws.gbuf);
  T* hbuf = ub ? static_cast<T*>(ub[7].h) : reinterpret_cast<T*>(wsp + ws.hbuf);
  T* stpbuf = reinterpret_cast<T*>(wsp + ws.stpbuf);
  T* feat = reinterpret_cast<T*>(wsp + ws.feat);
  T* fact = reinterpret_cast<T*>(wsp + ws.fact);
  T* h1 = reinterpret_cast<T*>(wsp + ws.h1);
  T* h2 = reinterpret_cast<T*>(wsp + ws.h2);
  float* params = reinterpret_cast<float*>(wsp + ws.params);
  const long long M = d.M(), hw = d.hw();
  const int xp = ctx->xpad3;
  const long long slabM = dense_slab(ctx, d);
  // LR ingest: x1 of the reversed block 8, and the X slot of local_m1
  bool ingested = false;
  if constexpr (kTc<T>) {
    if (slabM != 0 && xp == 16) {
      PROF(ctx, st, 5, (double)M * (12 + 16 + 3 * 16 * sizeof(T)), launch_lr_ingest_slab(lr, z, as_bf(gbuf), as_bf(hbuf), as_bf(stpbuf), M, hw, st, kX2<T>));
      ingested = true;
    }
  }
  if (!ingested) {
    PROF(ctx, st, 5, (double)M * 28, launch_nchw_to_dense<float>(lr, z, 4, 0, 0, 3, 4, M, hw, st));
    PROF(ctx, st, 5, (double)M * 28, launch_nchw_to_dense<T>(lr, gbuf, ws.gpitch, slabM, 0, 3, xp, M, hw, st));
    PROF(ctx, st, 5, (double)M * 28, launch_nchw_to_dense<T>(lr, hbuf, ws.gpitch, slabM, 0, 3, xp, M, hw, st));
    PROF(ctx, st, 5, (double)M * 28, launch_nchw_to_dense<T>(lr, stpbuf, ws.gpitch, slabM, 0, 3, xp, M, hw, st));
  }
  // STPNet.forward (:366-374)
  for (int i = 0; i < 6; ++i) {
    const DenseW& W = ctx->stp[i];
    const int pitch = i == 0 ? ws.gpitch : ws.spitch;
    SELFC_TRY(run_dense_convs<T>(ctx, W, stpbuf, pitch, d, st));
    ConvArgs<T> a = conv5_args<T>(ctx, W, stpbuf, pitch, d);
    a.epi = EPI_STORE; a.act = 0; a.outT = feat; a.outT_pitch = kStpC; a.outT_off = 0;
    PROF(ctx, st, 1, conv5_flops(W, d), launch_temporal<T>(ctx, W.t5, a, d, st));
    // training: every GlobalAgg output (= input of the next STP stage / the final feature) is also kept as fp32 [M][64]
    float* ga_keep = hooks && hooks->ga_save ? hooks->ga_save + (size_t)(i + 1) * d.M() * kStpC : nullptr;
    SELFC_TRY(run_global_agg<T>(ctx, ctx->ga[i], feat, stpbuf, ws.spitch, slabM, ga_keep, ga_keep ? kStpC : 0, nullptr, wsp, ws, d, st,
                                i == 5 ? fact : nullptr));
  }
  // tail_gmm (:336-344,:379): lrelu -> 64->128 -> lrelu -> 128->256 -> lrelu -> 256->720
  bool head_done = false, sampled = false, params_are_half = false;
  if constexpr (kTc<T>) {
    if (mode_tc(ctx) && ctx->head.t[0].img != nullptr) {
      // pointwise GEMMs over all M pixels, viewed as 2 pseudo-frames of ceil(M/2) rows (two accumulators in flight)
      // The 720-channel parameter tensor between the last head GEMM and the sampler is stored as fp16 quads: the GEMM's inputs are
      // bf16 (2^-9), fp16 (2^-11) adds nothing measurable, and its write + read halve.  SELFC_GMM_FP16=0: fp32 quads (also taken when
      // the thread-per-pixel sampler form is selected, which reads fp32 only).
      static int half_on = -1;
      if (half_on < 0) {
        const char* e = getenv("SELFC_GMM_FP16");
        const char* e2 = getenv("SELFC_GMM_SPLIT");
        half_on = ((e && atoi(e) == 0) || (e2 && atoi(e2) == 0)) ? 0 : 1;
      }
      const bool params_half = half_on == 1 && !kX2<T>;      // BF16X3 mode keeps fp32 parameters
      params_are_half = params_half;
      auto pointwise = [&](const TcTempW& w, const __nv_bfloat16* in, int in_pitch, __nv_bfloat16* outT, int outT_pitch, float* outF,
                           int outF_pitch, int outF_off, int act) -> int {
        TcTempArgs t;
        t.in = in; t.in_pitch = in_pitch; t.B = 1; t.T = 2; t.hw = (int)((M + 1) / 2); t.m_limit = M;
        t.epi = EPI_STORE; t.act = act;
        t.outT = outT; t.outT_pitch = outT_pitch; t.outF = outF; t.outF_pitch = outF_pitch; t.outF_off = outF_off;
        t.outF_planar = outF != nullptr ? (params_half ? 2 : 1) : 0;     // GMM parameters as planar quads (fp16 by default) for the sampler
        return launch_temporal_tc(w, t, st);
      };
      PROF(ctx, st, 3, 2.0 * M * 64 * 128, pointwise(ctx->head.t[0], as_bf(fact), kStpC, as_bf(h1), 128, nullptr, 0, 0, 1));
      PROF(ctx, st, 3, 2.0 * M * 128 * 256, pointwise(ctx->head.t[1], as_bf(h1), 128, as_bf(h2), 256, nullptr, 0, 0, 1));
      // SELFC_GMM_FUSED=1: head GEMM with the sampler fused into its epilogue, one launch per mixture component (the
      // 720-channel tensor never exists: 2.6 GB less workspace traffic at 1080p).  Parity-tested, but OFF by default: the
      // Philox / Box-Muller / exp work then runs on the GEMM's four epilogue warps per SM and is latency-bound there
      // (5 x 0.61 ms vs 3 x 0.245 ms GEMM + 0.98 ms full-occupancy sampler, measured).
      static int fused = -1;
      if (fused < 0) {
        const char* e = getenv("SELFC_GMM_FUSED");
        fused = (e && atoi(e) != 0) ? 1 : 0;
      }
      if (kX2<T>) {
        // (hi, lo) weights of a 256 -> 240 GEMM do not fit shared memory: one 256 -> 144 launch per mixture component (rows
        // [logit | log-scale | mean] x 48 of component k), written into the planar parameter tensor at channels j*240 + k*48 + hf
        for (int k = 0; k < kGmmK; ++k) {
          TcTempArgs t;
          t.in = as_bf(h2); t.in_pitch = 256; t.B = 1; t.T = 2; t.hw = (int)((M + 1) / 2); t.m_limit = M;
          t.epi = EPI_STORE; t.act = 0;
          t.outF = params; t.outF_pitch = 720; t.outF_off = kHF * k; t.outF_planar = 1; t.outF_blk = kHF; t.outF_blk_stride = 240;
          PROF(ctx, st, 3, 2.0 * M * 256 * 144, launch_temporal_tc(ctx->head.g[k], t, st));
        }
      } else if (fused && ctx->head.g[0].img != nullptr) {
        // 256 -> 720 and the soft-GMM draw, one launch per mixture component: the 720-channel tensor never exists
        for (int k = 0; k < kGmmK; ++k) {
          TcTempArgs t;
          t.in = as_bf(h2); t.in_pitch = 256; t.B = 1; t.T = 2; t.hw = (int)((M + 1) / 2); t.m_limit = M;
          t.epi = EPI_GMM; t.z = z;
          t.gmm_k = k; t.gmm_T = d.T; t.gmm_hw = hw; t.eps = eps; t.seed = seed; t.offset = offset;
          PROF(ctx, st, k == 0 ? 3 : 4, k == 0 ? 2.0 * M * 256 * 720 : (double)M * (720 + 48) * 4, launch_temporal_tc(ctx->head.g[k], t, st));
        }
        sampled = true;
      } else {
        for (int j = 0; j < 3; ++j)
          PROF(ctx, st, 3, 2.0 * M * 256 * 240, pointwise(ctx->head.t[2 + j], as_bf(h2), 256, nullptr, 0, params, 720, 240 * j, 0));
      }
      head_done = true;
    }
  }
  if (!head_done) {
    ConvArgs<T> a;
    a.BT = d.B * d.T; a.Tn = d.T; a.h = d.h; a.w_ = d.w;
    a.taps = 1; a.tap_mode = TAP_POINT; a.epi = EPI_STORE;
    a.in = stpbuf; a.in_pitch = ws.spitch; a.cin = 64; a.in_lrelu = 1; a.in_slabM = slabM;
    a.w = ctx->head.w[0]; a.bias = ctx->head.b[0]; a.np = ctx->head.np[0]; a.cout = 128;
    a.act = 1; a.outT = h1; a.outT_pitch = 128;
    PROF(ctx, st, 3, 2.0 * M * 64 * 128, launch_conv_simt<T>(a, st));
    a.in = h1; a.in_pitch = 128; a.cin = 128; a.in_lrelu = 0; a.in_slabM = 0;
    a.w = ctx->head.w[1]; a.bias = ctx->head.b[1]; a.np = ctx->head.np[1]; a.cout = 256;
    a.outT = h2; a.outT_pitch = 256;
    PROF(ctx, st, 3, 2.0 * M * 128 * 256, launch_conv_simt<T>(a, st));
    a.in = h2; a.in_pitch = 256; a.cin = 256;
    a.w = ctx->head.w[2]; a.bias = ctx->head.b[2]; a.np = ctx->head.np[2]; a.cout = 720;
    a.act = 0; a.outT = nullptr; a.outF = params; a.outF_pitch = 720;
    PROF(ctx, st, 3, 2.0 * M * 256 * 720, launch_conv_simt<T>(a, st));
  }
  if (sampled) {
  } else if (head_done)
    PROF(ctx, st, 4, (double)M * (720 * (params_are_half ? 2 : 4) + 48 * 4),
         launch_gmm_sample_planar(params, eps, seed, offset, z, d.B, d.T, d.h, d.w, st, -1, params_are_half));
  else
    PROF(ctx, st, 4, (double)M * (720 + 48) * 4, launch_gmm_sample(params, false, eps, seed, offset, z, false, /*planar z*/ -1, 0, d.B, d.T, d.h, d.w, st));
  if (hf) PROF(ctx, st, 5, (double)M * 48 * 8, launch_export_hf(z, hf, M, hw, st));
  for (int blk = 7; blk >= 0; --blk) {
    if (hooks && hooks->z_save)      // training: the state each reverse block starts from
      SELFC_CUDA(cudaMemcpyAsync(hooks->z_save + (size_t)blk * d.M() * kZQuads * 4, z, (size_t)d.M() * kZQuads * 16, cudaMemcpyDeviceToDevice, st));
    SELFC_TRY(run_invblock<T>(ctx, blk, true, wsp, ws, d, st, ub ? &ub[blk] : nullptr));
  }
  if (hooks && hooks->z_save)
    SELFC_CUDA(cudaMemcpyAsync(hooks->z_save + (size_t)8 * d.M() * kZQuads * 4, z, (size_t)d.M() * kZQuads * 16, cudaMemcpyDeviceToDevice, st));
  if (hr_img != nullptr)
    PROF(ctx, st, 5, (double)M * (51 * 4 + 48), launch_fa_rev_u8(z, hr_img, d.B * d.T, d.h, d.w, st));
  else
    PROF(ctx, st, 5, (double)M * (51 * 4 + 48 * 4), launch_fa_rev(z, false, hr, d.B * d.T, d.h, d.w, st));
  return 0;
}

// ---- weight packing ---------------------------------------------------------------------------------------
struct ArenaPlan {
  size_t off = 0;
  size_t take(size_t bytes) {
    size_t o = off;
    off = align_up(off + bytes, 256);
    return o;
  }
};

static void plan_dense(ArenaPlan& pl, int cin, int cout, int xpad, size_t woff[5], size_t boff[5], int np[5]) {
  for (int k = 0; k < 5; ++k) {
    const int taps = k < 4 ? 9 : 3;
    const int cin_buf = xpad + kGrowth * k;
    np[k] = k < 4 ? 32 : (int)align_up(cout, 32);
    woff[k] = pl.take((size_t)taps * cin_buf * np[k] * 4);
    boff[k] = pl.take((size_t)np[k] * 4);
  }
  (void)cin;
}

template <typename T>
static int d2dt_impl(selfc_ctx* ctx, const DenseW& W, const float* x, float* y, const Dims& d, char* wsp, const Workspace& ws,
                     cudaStream_t st) {
  T* buf = reinterpret_cast<T*>(wsp + ws.stpbuf);
  float* tmp = reinterpret_cast<float*>(wsp + ws.params);
  const int pitch = W.xpad + 4 * kGrowth;
  SELFC_TRY(launch_nchw_to_dense<T>(x, buf, pitch, dense_slab(ctx, d), 0, W.cin, W.xpad, d.M(), d.hw(), st));
  SELFC_TRY(run_dense_convs<T>(ctx, W, buf, pitch, d, st));
  ConvArgs<T> a = conv5_args<T>(ctx, W, buf, pitch, d);
  a.epi = EPI_STORE; a.outF = tmp; a.outF_pitch = 64;
  SELFC_TRY(launch_temporal<T>(ctx, W.t5, a, d, st));
  return launch_dense_to_nchw<float>(tmp, 64, 0, 0, y, W.cout, d.M(), d.hw(), st);
}

template <typename T>
static int conv3x3_impl(selfc_ctx* ctx, const DenseW& W, int k, const float* x, float* y, const Dims& d, char* wsp, const Workspace& ws,
                        cudaStream_t st) {
  T* buf = reinterpret_cast<T*>(wsp + ws.stpbuf);
  const int pitch = W.xpad + 4 * kGrowth;
  const int cref = W.cin + kGrowth * k;
  const long long slabM = dense_slab(ctx, d);
  SELFC_TRY(launch_nchw_slice_to_dense<T>(x, cref, 0, buf, pitch, slabM, 0, W.cin, W.xpad, d.M(), d.hw(), st));
  if (k > 0)
    SELFC_TRY(launch_nchw_slice_to_dense<T>(x, cref, W.cin, buf, pitch, slabM, W.xpad, kGrowth * k, kGrowth * k, d.M(), d.hw(), st));
  SELFC_TRY(run_dense_convs<T>(ctx, W, buf, pitch, d, st, k, k));
  return launch_dense_to_nchw<T>(buf, pitch, slabM, W.xpad + kGrowth * k, y, kGrowth, d.M(), d.hw(), st);
}

template <typename T>
static int ga_impl(selfc_ctx* ctx, const GaW& g, const float* x, float* y, float* wmat_out, const Dims& d, char* wsp,
                   const Workspace& ws, cudaStream_t st) {
  T* feat = reinterpret_cast<T*>(wsp + ws.feat);
  float* tmp = reinterpret_cast<float*>(wsp + ws.params);
  SELFC_TRY(launch_nchw_to_dense<T>(x, feat, kStpC, 0, 0, kStpC, kStpC, d.M(), d.hw(), st));
  SELFC_TRY(run_global_agg<T>(ctx, g, feat, nullptr, 0, 0, tmp, 64, wmat_out, wsp, ws, d, st));
  return launch_dense_to_nchw<float>(tmp, 64, 0, 0, y, kStpC, d.M(), d.hw(), st);
}

}  // namespace selfc

namespace selfc {
int check_run(selfc_ctx* ctx, int B, int T, int H, int W, void* workspace, size_t workspace_bytes, Workspace* ws) {
  SELFC_CHECK_ARG(ctx != nullptr, "null context");
  if (!ctx->loaded) {
    set_error("weights not loaded: call selfc_ctx_load_weights first");
    return SELFC_E_STATE;
  }
  SELFC_CHECK_ARG(B >= 1 && T >= 1 && T <= 32, "B=%d T=%d: need B >= 1 and 1 <= T <= 32", B, T);
  SELFC_CHECK_ARG(H >= 4 && W >= 4 && H % 4 == 0 && W % 4 == 0, "H=%d W=%d must be positive multiples of 4", H, W);
  SELFC_CHECK_ARG((long long)B * T * (H / 4) * (W / 4) < (1ll << 31) / 8, "clip batch too large for 32-bit tile indices");
  *ws = make_workspace(ctx, B, T, H / 4, W / 4);
  SELFC_CHECK_ARG(workspace != nullptr && aligned16(workspace), "workspace null or misaligned");
  // BF16X3 mode addresses the two halves of an element from the low bits of its handle (common.cuh: 64-byte rows)
  SELFC_CHECK_ARG(ctx->mode != SELFC_MODE_BF16X3 || (reinterpret_cast<uintptr_t>(workspace) & 63u) == 0, "BF16X3 mode needs a 64-byte aligned workspace");
  if (workspace_bytes < ws->total) {
    set_error("workspace too small: %zu < %zu bytes", workspace_bytes, ws->total);
    return SELFC_E_STATE;
  }
  SELFC_CUDA(cudaSetDevice(ctx->device));
  return 0;
}

}  // namespace selfc

namespace selfc {
const DenseW* find_dense(selfc_ctx* ctx, int first_param) {
  if (first_param >= 0 && first_param < 240 && first_param % 10 == 0) return &ctx->inv[first_param / 30][(first_param % 30) / 10];
  if (first_param == P_LOCAL1) return &ctx->stp[0];
  if (first_param == P_LOCAL2) return &ctx->stp[1];
  for (int i = 0; i < 4; ++i)
    if (first_param == P_OTHER + 18 * i) return &ctx->stp[2 + i];
  return nullptr;
}

template <typename E>
int invblock_fwd(const selfc_ctx* ctx, int blk, bool rev, char* wsp, const Workspace& ws, const Dims& d, cudaStream_t st, const BlockBufsV* bufs) {
  return run_invblock<E>(ctx, blk, rev, wsp, ws, d, st, bufs);
}
template <typename E>
int up_hooked(selfc_ctx* ctx, const float* lr, const float* eps, uint64_t seed, uint64_t offset, float* hr, const Dims& d, char* wsp,
              const Workspace& ws, cudaStream_t st, const TrainHooks* hooks) {
  return up_impl<E>(ctx, lr, eps, seed, offset, hr, nullptr, d, wsp, ws, st, hooks);
}
template <typename E>
int stp_dense(const selfc_ctx* ctx, int i, E* stpbuf, int pitch, float* feat, const Dims& d, cudaStream_t st) {
  const DenseW& W = ctx->stp[i];
  SELFC_TRY(run_dense_convs<E>(ctx, W, stpbuf, pitch, d, st));
  ConvArgs<E> a = conv5_args<E>(ctx, W, stpbuf, pitch, d);
  a.epi = EPI_STORE; a.act = 0; a.outF = feat; a.outF_pitch = kStpC; a.outF_off = 0;
  return launch_temporal<E>(ctx, W.t5, a, d, st);
}
template <typename E>
int dense_convs(const selfc_ctx* ctx, const DenseW& W, E* buf, int pitch, const Dims& d, cudaStream_t st) {
  return run_dense_convs<E>(ctx, W, buf, pitch, d, st);
}
template int invblock_fwd<float>(const selfc_ctx*, int, bool, char*, const Workspace&, const Dims&, cudaStream_t, const BlockBufsV*);
template int invblock_fwd<bfx2>(const selfc_ctx*, int, bool, char*, const Workspace&, const Dims&, cudaStream_t, const BlockBufsV*);
template int up_hooked<float>(selfc_ctx*, const float*, const float*, uint64_t, uint64_t, float*, const Dims&, char*, const Workspace&, cudaStream_t,
                              const TrainHooks*);
template int up_hooked<bfx2>(selfc_ctx*, const float*, const float*, uint64_t, uint64_t, float*, const Dims&, char*, const Workspace&, cudaStream_t,
                             const TrainHooks*);
template int stp_dense<float>(const selfc_ctx*, int, float*, int, float*, const Dims&, cudaStream_t);
template int stp_dense<bfx2>(const selfc_ctx*, int, bfx2*, int, float*, const Dims&, cudaStream_t);
template int dense_convs<float>(const selfc_ctx*, const DenseW&, float*, int, const Dims&, cudaStream_t);
template int dense_convs<bfx2>(const selfc_ctx*, const DenseW&, bfx2*, int, const Dims&, cudaStream_t);
const GaW* find_ga(selfc_ctx* ctx, int first_param) {
  const int ga_first[6] = {P_GLOBAL1, P_GLOBAL2, P_OTHER + 10, P_OTHER + 28, P_OTHER + 46, P_OTHER + 64};
  for (int i = 0; i < 6; ++i)
    if (first_param == ga_first[i]) return &ctx->ga[i];
  return nullptr;
}
}  // namespace selfc

// =============================================================================================================
// C-ABI
// =============================================================================================================
namespace selfc {
PackBatch*& pack_batch_current() {
  static thread_local PackBatch* cur = nullptr;
  return cur;
}
int pack_batch_flush(PackBatch& b, cudaStream_t st) {
  SELFC_CHECK_ARG(b.point < kPackFlushPoints, "pack batch: more than %d flush points", kPackFlushPoints);
  SELFC_TRY(flush_pack_conv_simt(b.simt[b.point], st));
  SELFC_TRY(flush_pack_tc3(b.tc3[b.point], st));
  SELFC_TRY(flush_pack_dgrad_slot(b.slot[b.point], st));
  SELFC_TRY(flush_pack_dgrad5_ref(b.ref5[b.point], st));
  SELFC_TRY(flush_pack_temporal(b.temporal[b.point], st));
  ++b.point;
  return 0;
}
bool pack_batch_enabled() {
  const char* e = getenv("SELFC_PACK_BATCH");
  return !(e && atoi(e) == 0);
}
}  // namespace selfc

extern "C" {

int selfc_version(void) { return 100; }
int selfc_dense_fused_schedule(int sch, int* out18) { return selfc::dense_fused_schedule(sch, out18); }
/* the zero-padded pixel planes of the tensor-core weight-gradient kernel (wgrad_tc.cu) for a clip batch: out4 = {Wp, Fp, P, Pa};
 * pixel (b, t, y, x) lives at P = ((b * (T + 1) + t + 1) * (h + 2) + y + 1) * Wp + x + 1 (host-side check of the padding scheme) */
int selfc_wgrad_geometry(int B, int T, int h, int w, long long* out4) {
  SELFC_CHECK_ARG(B >= 1 && T >= 1 && h >= 1 && w >= 1 && out4 != nullptr, "wgrad_geometry: bad argument");
  const selfc::WgGeom g = selfc::wg_geometry(selfc::Dims{B, T, h, w});
  out4[0] = g.Wp; out4[1] = g.Fp; out4[2] = g.P; out4[3] = g.Pa;
  return 0;
}
/* debug only (SELFC_TC_DBG=1): barrier-wait cycle counters of the tcgen05 temporal kernel, 17 int64 per launch */
int selfc_debug_read(long long* out, int cap) { return selfc::tc::debug_read(out, cap); }
const char* selfc_last_error(void) { return g_err; }
uint64_t selfc_launch_count(void) { return g_launches; }

int selfc_ctx_create(selfc_ctx** out, int device, int mode) {
  SELFC_CHECK_ARG(out != nullptr, "selfc_ctx_create: null out");
  SELFC_CHECK_ARG(mode == SELFC_MODE_FP32 || mode == SELFC_MODE_BF16 || mode == SELFC_MODE_BF16X3, "selfc_ctx_create: unknown mode %d", mode);
  int ndev = 0;
  SELFC_CUDA(cudaGetDeviceCount(&ndev));
  SELFC_CHECK_ARG(device >= 0 && device < ndev, "selfc_ctx_create: device %d of %d", device, ndev);
  cudaDeviceProp prop;
  SELFC_CUDA(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10) {
    set_error("selfc_b200 is built for sm_100a only; device %d is sm_%d%d", device, prop.major, prop.minor);
    return SELFC_E_UNSUPPORTED;
  }
  selfc_ctx* c = new (std::nothrow) selfc_ctx();
  SELFC_CHECK_ARG(c != nullptr, "out of host memory");
  c->device = device;
  c->mode = mode;
  c->xpad3 = mode == SELFC_MODE_FP32 ? 4 : 16;
  *out = c;
  return 0;
}

int selfc_ctx_destroy(selfc_ctx* ctx) {
  if (!ctx) return 0;
  if (ctx->arena) cudaFree(ctx->arena);
  if (ctx->train_scratch) cudaFree(ctx->train_scratch);
  if (ctx->train_zero_bias) cudaFree(ctx->train_zero_bias);
  if (ctx->wg_planes) cudaFree(ctx->wg_planes);
  if (ctx->dg_gslab) cudaFree(ctx->dg_gslab);
  ctx->pack.release();
  ctx->pack_dg.release();
  auto free_dg = [](DenseW& W) {
    if (W.dg5_tmp) cudaFree(W.dg5_tmp);
    for (int k = 0; k < 4; ++k)
      if (W.dg_img[k]) cudaFree(W.dg_img[k]);
    free_temporal_weights(W.dg5[0]);
    free_temporal_weights(W.dg5[1]);
  };
  for (int b = 0; b < 8; ++b)
    for (int j = 0; j < 3; ++j) free_dg(ctx->inv[b][j]);
  for (int i = 0; i < 6; ++i) free_dg(ctx->stp[i]);
  for (int b = 0; b < 8; ++b)
    for (int j = 0; j < 3; ++j)
    {
      for (int k = 0; k < 5; ++k) free_tc_weights(ctx->inv[b][j].tc[k]);
      free_temporal_weights(ctx->inv[b][j].t5);
      if (ctx->inv[b][j].f5img) cudaFree(ctx->inv[b][j].f5img);
    }
  for (int i = 0; i < 6; ++i) {
    for (int k = 0; k < 5; ++k) free_tc_weights(ctx->stp[i].tc[k]);
    free_temporal_weights(ctx->stp[i].t5);
    free_temporal_weights(ctx->ga[i].tp);
    if (ctx->ga[i].wmap_cache) cudaFree(ctx->ga[i].wmap_cache);
  }
  for (int j = 0; j < 5; ++j) free_temporal_weights(ctx->head.t[j]);
  for (int j = 0; j < 5; ++j) free_temporal_weights(ctx->head.g[j]);
  for (int j = 0; j < 5; ++j) free_temporal_weights(ctx->head.r[j]);
  for (int j = 0; j < 4; ++j) free_temporal_weights(ctx->head.dg3[j]);
  free_temporal_weights(ctx->head.dg2);
  free_temporal_weights(ctx->head.dg1);
  delete ctx;
  return 0;
}

int selfc_ctx_mode(const selfc_ctx* ctx) { return ctx ? ctx->mode : SELFC_E_ARG; }

int selfc_prof_enable(selfc_ctx* ctx, int on) {
  SELFC_CHECK_ARG(ctx != nullptr, "null context");
  for (auto& r : ctx->prof) {
    cudaEventDestroy(r.a);
    cudaEventDestroy(r.b);
  }
  ctx->prof.clear();
  ctx->prof_on = on != 0;
  return 0;
}

int selfc_prof_read(selfc_ctx* ctx, int ncls, double* ms, double* work, uint64_t* launches) {
  SELFC_CHECK_ARG(ctx && ms && work && launches && ncls >= 1, "prof_read: bad argument");
  for (int i = 0; i < ncls; ++i) { ms[i] = 0.0; work[i] = 0.0; launches[i] = 0; }
  for (auto& r : ctx->prof) {
    SELFC_CUDA(cudaEventSynchronize(r.b));
    float t = 0.f;
    SELFC_CUDA(cudaEventElapsedTime(&t, r.a, r.b));
    if (r.cls >= 0 && r.cls < ncls) { ms[r.cls] += t; work[r.cls] += r.work; launches[r.cls] += 1; }
    cudaEventDestroy(r.a);
    cudaEventDestroy(r.b);
  }
  ctx->prof.clear();
  return 0;
}

int selfc_ctx_load_weights(selfc_ctx* ctx, const float* const* p, int n_params, void* stream) {
  SELFC_CHECK_ARG(ctx && p, "selfc_ctx_load_weights: null argument");
  SELFC_CHECK_ARG(n_params == SELFC_NUM_PARAMS, "selfc_ctx_load_weights: expected %d tensors, got %d", SELFC_NUM_PARAMS, n_params);
  for (int i = 0; i < n_params; ++i) SELFC_CHECK_ARG(p[i] != nullptr, "selfc_ctx_load_weights: tensor %d is null", i);
  cudaStream_t st = (cudaStream_t)stream;
  std::lock_guard<std::mutex> lock(ctx->mu);
  SELFC_CUDA(cudaSetDevice(ctx->device));
  const int xp3 = ctx->xpad3;
  const bool x2 = ctx->mode == SELFC_MODE_BF16X3;
  // the ~300 pack launches below are recorded and run as three launches per flush point (pack_batch.h).  A flush precedes every launch
  // that overwrites something a recorded pack reads (the permuted head weights)
  std::unique_ptr<PackBatchScope> batch_scope;
  if (pack_batch_enabled()) batch_scope.reset(new PackBatchScope(ctx->pack));
  auto flush_packs = [&]() -> int { return batch_scope ? pack_batch_flush(ctx->pack, st) : 0; };

  // plan the arena
  ArenaPlan pl;
  struct DensePlan { size_t w[5], b[5]; int np[5]; int cin, cout, xpad, first; };
  std::vector<DensePlan> dp;
  auto add_dense = [&](int first, int cin, int cout) {
    DensePlan d;
    d.cin = cin; d.cout = cout; d.first = first;
    d.xpad = cin == 3 ? xp3 : cin;
    plan_dense(pl, cin, cout, d.xpad, d.w, d.b, d.np);
    dp.push_back(d);
  };
  for (int b = 0; b < 8; ++b) {
    add_dense(P_INV0 + b * 30 + 0, kHF, 3);     // F
    add_dense(P_INV0 + b * 30 + 10, 3, kHF);    // G
    add_dense(P_INV0 + b * 30 + 20, 3, kHF);    // H
  }
  add_dense(P_LOCAL1, 3, kStpC);
  add_dense(P_LOCAL2, kStpC, kStpC);
  for (int i = 0; i < 4; ++i) add_dense(P_OTHER + i * 18, kStpC, kStpC);
  const int ga_first[6] = {P_GLOBAL1, P_GLOBAL2, P_OTHER + 10, P_OTHER + 28, P_OTHER + 46, P_OTHER + 64};
  size_t ga_off[6][8];
  const size_t ga_sz[8] = {1024, 1, 64 * 64, 64, 64 * 64, 64, 64 * 64, 64};   // fc.w fc.b p1.w p1.b p2.w p2.b p3.w p3.b
  for (int g = 0; g < 6; ++g)
    for (int j = 0; j < 8; ++j) ga_off[g][j] = pl.take(ga_sz[j] * 4);
  size_t head_w[3], head_b[3];
  for (int j = 0; j < 3; ++j) {
    head_w[j] = pl.take((size_t)ctx->head.cin[j] * ctx->head.np[j] * 4);
    head_b[j] = pl.take((size_t)ctx->head.np[j] * 4);
  }
  const size_t head_perm_w = pl.take((size_t)720 * 256 * 4);
  const size_t head_perm_b = pl.take(720 * 4);
  const size_t head_zero = pl.take(256 * 4);
  if (ctx->arena_bytes < pl.off) {
    if (ctx->arena) cudaFree(ctx->arena);
  if (ctx->train_scratch) cudaFree(ctx->train_scratch);
  if (ctx->train_zero_bias) cudaFree(ctx->train_zero_bias);
    ctx->arena = nullptr;
    SELFC_CUDA(cudaMalloc(&ctx->arena, pl.off));
    ctx->arena_bytes = pl.off;
  }
  char* A = ctx->arena;
  auto fp = [&](size_t off) { return reinterpret_cast<float*>(A + off); };

  // dense blocks
  for (size_t i = 0; i < dp.size(); ++i) {
    const DensePlan& d = dp[i];
    DenseW* W = i < 24 ? &ctx->inv[i / 3][i % 3] : &ctx->stp[i - 24];
    W->cin = d.cin; W->cout = d.cout; W->xpad = d.xpad;
    W->dg_valid = false;            // the input-gradient images follow the weights: re-packed by the next backward
    for (int k = 0; k < 5; ++k) {
      const int taps = k < 4 ? 9 : 3;
      const int cin_ref = d.cin + kGrowth * k;
      const int cin_buf = d.xpad + kGrowth * k;
      const int cout = k < 4 ? kGrowth : d.cout;
      W->w[k] = fp(d.w[k]); W->b[k] = fp(d.b[k]); W->np[k] = d.np[k];
      SELFC_TRY(launch_pack_conv_simt(p[d.first + 2 * k], p[d.first + 2 * k + 1], W->w[k], W->b[k], cout, cin_ref, taps, cin_buf,
                                      d.cin, d.xpad, d.np[k], st));
      if (mode_tc(ctx) && k < 4) {
        SELFC_TRY(pack_tc_weights(W->tc[k], p[d.first + 2 * k], p[d.first + 2 * k + 1], cin_ref, cin_buf, d.cin, d.xpad, st, x2));
      }
      if (mode_tc(ctx) && k == 4) {
        SELFC_TRY(pack_temporal_weights(W->t5, p[d.first + 8], p[d.first + 9], d.cout, cin_ref, 3, cin_buf, d.cin, d.xpad, st, x2));
        if (!x2 && d.cout == 3 && d.cin == d.xpad && cin_buf == 176) SELFC_TRY(pack_f5_weights(&W->f5img, p[d.first + 8], cin_buf, st));
      }
    }
  }
  // GlobalAgg
  for (int g = 0; g < 6; ++g) {
    GaW& G = ctx->ga[g];
    G.wmap_h = G.wmap_w = 0;          // fc.weight changes: the cached pooling weight map is stale
    const int f = ga_first[g];
    float* dst[8];
    for (int j = 0; j < 8; ++j) dst[j] = fp(ga_off[g][j]);
    G.fcw = dst[0]; G.fcb = dst[1]; G.p1w = dst[2]; G.p1b = dst[3]; G.p2w = dst[4]; G.p2b = dst[5]; G.p3w = dst[6]; G.p3b = dst[7];
    SELFC_CUDA(cudaMemcpyAsync(G.fcw, p[f + 0], 1024 * 4, cudaMemcpyDeviceToDevice, st));
    SELFC_CUDA(cudaMemcpyAsync(G.fcb, p[f + 1], 4, cudaMemcpyDeviceToDevice, st));
    SELFC_TRY(launch_pack_conv_simt(p[f + 2], p[f + 3], G.p1w, G.p1b, 64, 64, 1, 64, 64, 64, 64, st));
    if (mode_tc(ctx)) SELFC_TRY(pack_temporal_weights(G.tp, p[f + 2], p[f + 3], 64, 64, 1, 64, 64, 64, st, x2));
    SELFC_CUDA(cudaMemcpyAsync(G.p2w, p[f + 4], 64 * 64 * 4, cudaMemcpyDeviceToDevice, st));
    SELFC_CUDA(cudaMemcpyAsync(G.p2b, p[f + 5], 64 * 4, cudaMemcpyDeviceToDevice, st));
    SELFC_CUDA(cudaMemcpyAsync(G.p3w, p[f + 6], 64 * 64 * 4, cudaMemcpyDeviceToDevice, st));
    SELFC_CUDA(cudaMemcpyAsync(G.p3b, p[f + 7], 64 * 4, cudaMemcpyDeviceToDevice, st));
  }
  // GMM head
  for (int j = 0; j < 3; ++j) {
    ctx->head.w[j] = fp(head_w[j]);
    ctx->head.b[j] = fp(head_b[j]);
    const int cin = ctx->head.cin[j], cout = ctx->head.cout[j];
    SELFC_TRY(launch_pack_conv_simt(p[P_TAIL + 2 * j], p[P_TAIL + 2 * j + 1], ctx->head.w[j], ctx->head.b[j], cout, cin, 1, cin, cin,
                                    cin, ctx->head.np[j], st));
  }
  if (mode_tc(ctx)) {
    SELFC_TRY(pack_temporal_weights(ctx->head.t[0], p[P_TAIL + 0], p[P_TAIL + 1], 128, 64, 1, 64, 64, 64, st, x2));
    SELFC_TRY(pack_temporal_weights(ctx->head.t[1], p[P_TAIL + 2], p[P_TAIL + 3], 256, 128, 1, 128, 128, 128, st, x2));
    float* wperm = fp(head_perm_w);
    float* bperm = fp(head_perm_b);
    if (!x2) {      // (hi, lo) images of a 256 -> 240 GEMM would not fit shared memory: BF16X3 mode uses the per-component images
      SELFC_TRY(launch_permute_gmm_rows(p[P_TAIL + 4], p[P_TAIL + 5], wperm, bperm, false, st));
      for (int j = 0; j < 3; ++j)
        SELFC_TRY(pack_temporal_weights(ctx->head.t[2 + j], wperm + (size_t)240 * j * 256, bperm + 240 * j, 240, 256, 1, 256, 256, 256, st));
    }
    SELFC_TRY(flush_packs());        // (the packs above read the first permutation)
    SELFC_TRY(launch_permute_gmm_rows(p[P_TAIL + 4], p[P_TAIL + 5], wperm, bperm, true, st));
    for (int k = 0; k < kGmmK; ++k)
      SELFC_TRY(pack_temporal_weights(ctx->head.g[k], wperm + (size_t)144 * k * 256, bperm + 144 * k, 144, 256, 1, 256, 256, 256, st, x2));
    if (x2) {
      // training (BF16X3): the head's last layer in reference channel order, and the three layers' input-gradient images.  The latter
      // come straight from the fp32-FMA packs [cin][np] (row = input channel = an input-gradient OUTPUT, column = output channel = its K)
      for (int i = 0; i < 5; ++i)
        SELFC_TRY(pack_temporal_weights(ctx->head.r[i], p[P_TAIL + 4] + (size_t)144 * i * 256, p[P_TAIL + 5] + 144 * i, 144, 256, 1, 256, 256, 256, st, true));
      // zero bias: the first 256 floats of the permuted-bias scratch are overwritten below, so use a dedicated zero vector in the arena
      float* zero = fp(head_zero);
      SELFC_CUDA(cudaMemsetAsync(zero, 0, 256 * sizeof(float), st));
      for (int j = 0; j < 4; ++j)
        SELFC_TRY(pack_temporal_weights(ctx->head.dg3[j], ctx->head.w[2] + (size_t)64 * j * ctx->head.np[2], zero, 64, ctx->head.np[2], 1, 720, 720, 720, st, true));
      SELFC_TRY(pack_temporal_weights(ctx->head.dg2, ctx->head.w[1], zero, 128, ctx->head.np[1], 1, 256, 256, 256, st, true));
      SELFC_TRY(pack_temporal_weights(ctx->head.dg1, ctx->head.w[0], zero, 64, ctx->head.np[0], 1, 128, 128, 128, st, true));
    }
  }
  SELFC_TRY(flush_packs());
  ctx->loaded = true;
  return 0;
}

size_t selfc_workspace_bytes(const selfc_ctx* ctx, int B, int T, int h, int w) {
  if (!ctx || B < 0 || T < 1 || h < 1 || w < 1) return 0;
  return make_workspace(ctx, B, T, h, w).total;
}

int selfc_down(selfc_ctx* ctx, const float* hr, float* out51, uint8_t* lr_u8, float* lr_q, int B, int T, int H, int W,
               void* workspace, size_t workspace_bytes, void* stream) {
  Workspace ws;
  SELFC_TRY(check_run(ctx, B, T, H, W, workspace, workspace_bytes, &ws));
  SELFC_CHECK_ARG(hr && aligned16(hr), "hr null or not 16-byte aligned");
  Dims d{B, T, H / 4, W / 4};
  cudaStream_t st = (cudaStream_t)stream;
  if (ctx->mode == SELFC_MODE_BF16)
    return down_impl<__nv_bfloat16>(ctx, hr, out51, lr_u8, lr_q, d, (char*)workspace, ws, st);
  if (ctx->mode == SELFC_MODE_BF16X3) return down_impl<bfx2>(ctx, hr, out51, lr_u8, lr_q, d, (char*)workspace, ws, st);
  return down_impl<float>(ctx, hr, out51, lr_u8, lr_q, d, (char*)workspace, ws, st);
}

int selfc_up(selfc_ctx* ctx, const float* lr, const float* eps, uint64_t seed, uint64_t offset, float* hr, float* hf, int B, int T,
             int H, int W, void* workspace, size_t workspace_bytes, void* stream) {
  Workspace ws;
  SELFC_TRY(check_run(ctx, B, T, H, W, workspace, workspace_bytes, &ws));
  SELFC_CHECK_ARG(lr && hr && aligned16(hr), "lr/hr null or hr not 16-byte aligned");
  Dims d{B, T, H / 4, W / 4};
  cudaStream_t st = (cudaStream_t)stream;
  if (ctx->mode == SELFC_MODE_BF16)
    return up_impl<__nv_bfloat16>(ctx, lr, eps, seed, offset, hr, hf, d, (char*)workspace, ws, st);
  if (ctx->mode == SELFC_MODE_BF16X3) return up_impl<bfx2>(ctx, lr, eps, seed, offset, hr, hf, d, (char*)workspace, ws, st);
  return up_impl<float>(ctx, lr, eps, seed, offset, hr, hf, d, (char*)workspace, ws, st);
}

// ---- 8-bit frames at the boundary (decoded PNG / raw video planes in cv2 layout [n][H][W][3], B,G,R) ----------------
int selfc_down_u8(selfc_ctx* ctx, const uint8_t* hr_img, uint8_t* lr_img, float* lr_q, int B, int T, int H, int W, void* workspace,
                  size_t workspace_bytes, void* stream) {
  Workspace ws;
  SELFC_TRY(check_run(ctx, B, T, H, W, workspace, workspace_bytes, &ws));
  SELFC_CHECK_ARG(hr_img && ((uintptr_t)hr_img & 3) == 0, "hr_img null or not 4-byte aligned");
  Dims d{B, T, H / 4, W / 4};
  cudaStream_t st = (cudaStream_t)stream;
  float* q = lr_q ? lr_q : reinterpret_cast<float*>((char*)workspace + ws.lrq);
  if (ctx->mode == SELFC_MODE_BF16)
    SELFC_TRY(down_impl<__nv_bfloat16>(ctx, nullptr, nullptr, nullptr, q, d, (char*)workspace, ws, st, hr_img));
  else if (ctx->mode == SELFC_MODE_BF16X3)
    SELFC_TRY(down_impl<bfx2>(ctx, nullptr, nullptr, nullptr, q, d, (char*)workspace, ws, st, hr_img));
  else
    SELFC_TRY(down_impl<float>(ctx, nullptr, nullptr, nullptr, q, d, (char*)workspace, ws, st, hr_img));
  if (lr_img) SELFC_TRY(launch_frames_to_u8(q, lr_img, (long long)B * T, d.hw(), st));
  return 0;
}

int selfc_up_u8(selfc_ctx* ctx, const uint8_t* lr_img, const float* eps, uint64_t seed, uint64_t offset, uint8_t* hr_img, int B, int T,
                int H, int W, void* workspace, size_t workspace_bytes, void* stream) {
  Workspace ws;
  SELFC_TRY(check_run(ctx, B, T, H, W, workspace, workspace_bytes, &ws));
  SELFC_CHECK_ARG(lr_img && hr_img && ((uintptr_t)hr_img & 3) == 0, "lr_img/hr_img null or hr_img not 4-byte aligned");
  Dims d{B, T, H / 4, W / 4};
  cudaStream_t st = (cudaStream_t)stream;
  float* q = reinterpret_cast<float*>((char*)workspace + ws.lrq);
  SELFC_TRY(launch_frames_from_u8(lr_img, q, (long long)B * T, d.hw(), st));
  if (ctx->mode == SELFC_MODE_BF16)
    return up_impl<__nv_bfloat16>(ctx, q, eps, seed, offset, nullptr, nullptr, d, (char*)workspace, ws, st, nullptr, hr_img);
  if (ctx->mode == SELFC_MODE_BF16X3) return up_impl<bfx2>(ctx, q, eps, seed, offset, nullptr, nullptr, d, (char*)workspace, ws, st, nullptr, hr_img);
  return up_impl<float>(ctx, q, eps, seed, offset, nullptr, nullptr, d, (char*)workspace, ws, st, nullptr, hr_img);
}

int selfc_rescale_u8(selfc_ctx* ctx, const uint8_t* hr_img, const float* eps, uint64_t seed, uint64_t offset, uint8_t* lr_img,
                     uint8_t* hr_out_img, int B, int T, int H, int W, void* workspace, size_t workspace_bytes, void* stream) {
  Workspace ws;
  SELFC_TRY(check_run(ctx, B, T, H, W, workspace, workspace_bytes, &ws));
  SELFC_CHECK_ARG(hr_img && hr_out_img && ((uintptr_t)hr_img & 3) == 0 && ((uintptr_t)hr_out_img & 3) == 0,
                  "hr_img/hr_out_img null or not 4-byte aligned");
  Dims d{B, T, H / 4, W / 4};
  cudaStream_t st = (cudaStream_t)stream;
  char* wsp = (char*)workspace;
  float* q = reinterpret_cast<float*>(wsp + ws.lrq);
  const bool bf = ctx->mode == SELFC_MODE_BF16, x3 = ctx->mode == SELFC_MODE_BF16X3;
  SELFC_TRY(bf ? down_impl<__nv_bfloat16>(ctx, nullptr, nullptr, nullptr, q, d, wsp, ws, st, hr_img)
               : x3 ? down_impl<bfx2>(ctx, nullptr, nullptr, nullptr, q, d, wsp, ws, st, hr_img)
                    : down_impl<float>(ctx, nullptr, nullptr, nullptr, q, d, wsp, ws, st, hr_img));
  if (lr_img) SELFC_TRY(launch_frames_to_u8(q, lr_img, (long long)B * T, d.hw(), st));
  return bf ? up_impl<__nv_bfloat16>(ctx, q, eps, seed, offset, nullptr, nullptr, d, wsp, ws, st, nullptr, hr_out_img)
            : x3 ? up_impl<bfx2>(ctx, q, eps, seed, offset, nullptr, nullptr, d, wsp, ws, st, nullptr, hr_out_img)
                 : up_impl<float>(ctx, q, eps, seed, offset, nullptr, nullptr, d, wsp, ws, st, nullptr, hr_out_img);
}

int selfc_frames_from_u8(const uint8_t* img, float* x, int N, int H, int W, void* stream) {
  SELFC_CHECK_ARG(N >= 0 && H >= 1 && W >= 1 && (N == 0 || (img && x)), "frames_from_u8: null pointer or N=%d H=%d W=%d", N, H, W);
  return launch_frames_from_u8(img, x, N, (long long)H * W, (cudaStream_t)stream);
}

int selfc_frames_to_u8(const float* x, uint8_t* img, int N, int H, int W, void* stream) {
  SELFC_CHECK_ARG(N >= 0 && H >= 1 && W >= 1 && (N == 0 || (img && x)), "frames_to_u8: null pointer or N=%d H=%d W=%d", N, H, W);
  return launch_frames_to_u8(x, img, N, (long long)H * W, (cudaStream_t)stream);
}

int selfc_fa_fwd(const float* x, float* out51, int N, int H, int W, void* stream) {
  SELFC_CHECK_ARG(x && out51 && aligned16(x), "fa_fwd: null or misaligned pointer");
  SELFC_CHECK_ARG(N >= 0 && H >= 4 && W >= 4 && H % 4 == 0 && W % 4 == 0, "fa_fwd: N=%d H=%d W=%d", N, H, W);
  if (N == 0) return 0;
  return launch_fa_fwd_nchw(x, out51, N, H / 4, W / 4, (cudaStream_t)stream);
}

int selfc_fa_rev(const float* z51, float* y, int N, int h, int w, void* stream) {
  SELFC_CHECK_ARG(z51 && y && aligned16(y), "fa_rev: null or misaligned pointer");
  SELFC_CHECK_ARG(N >= 0 && h >= 1 && w >= 1, "fa_rev: N=%d h=%d w=%d", N, h, w);
  if (N == 0) return 0;
  return launch_fa_rev(z51, true, y, N, h, w, (cudaStream_t)stream);
}

int selfc_rgb_to_y(const float* x, float* y, int N, int H, int W, void* stream) {
  SELFC_CHECK_ARG(N >= 0 && H >= 1 && W >= 1 && (N == 0 || (x && y)), "rgb_to_y: null pointer or N=%d H=%d W=%d", N, H, W);
  return launch_rgb_to_y(x, y, N, (long long)H * W, (cudaStream_t)stream);
}

int selfc_frame_metrics(const float* a, const float* b, int N, int C, int H, int W, int to_y, const float* win11, double* sse,
                        double* ssim_sum, void* stream) {
  SELFC_CHECK_ARG(N >= 0 && C >= 1 && H >= 1 && W >= 1, "frame_metrics: N=%d C=%d H=%d W=%d", N, C, H, W);
  if (N == 0) return 0;
  SELFC_CHECK_ARG(a && b && sse, "frame_metrics: null pointer");
  SELFC_CHECK_ARG(!to_y || C == 3, "frame_metrics: luma conversion needs 3 channels, got %d", C);
  SELFC_CHECK_ARG(ssim_sum == nullptr || (win11 != nullptr && H >= 11 && W >= 11), "frame_metrics: SSIM needs an 11-tap window and H,W >= 11");
  return launch_frame_metrics(a, b, N, C, H, W, to_y, win11, sse, ssim_sum, (cudaStream_t)stream);
}

int selfc_fa2_fwd(const float* x, float* out15, int N, int H, int W, void* stream) {
  SELFC_CHECK_ARG(N >= 0 && H >= 2 && W >= 2 && H % 2 == 0 && W % 2 == 0, "fa2_fwd: N=%d H=%d W=%d", N, H, W);
  if (N == 0) return 0;
  SELFC_CHECK_ARG(x && out15 && ((uintptr_t)x & 7) == 0, "fa2_fwd: null or misaligned pointer");
  return launch_fa2(x, out15, false, N, H / 2, W / 2, (cudaStream_t)stream);
}

int selfc_fa2_rev(const float* z15, float* y, int N, int h, int w, void* stream) {
  SELFC_CHECK_ARG(N >= 0 && h >= 1 && w >= 1, "fa2_rev: N=%d h=%d w=%d", N, h, w);
  if (N == 0) return 0;
  SELFC_CHECK_ARG(z15 && y && ((uintptr_t)y & 7) == 0, "fa2_rev: null or misaligned pointer");
  return launch_fa2(z15, y, true, N, h, w, (cudaStream_t)stream);
}

int selfc_haar_fwd(const float* x, float* out, int N, int C, int H, int W, void* stream) {
  SELFC_CHECK_ARG(N >= 0 && C >= 1 && H >= 2 && W >= 2 && H % 2 == 0 && W % 2 == 0, "haar_fwd: N=%d C=%d H=%d W=%d", N, C, H, W);
  if (N == 0) return 0;
  SELFC_CHECK_ARG(x && out && ((uintptr_t)x & 7) == 0, "haar_fwd: null or misaligned pointer");
  return launch_haar(x, out, false, N, C, H / 2, W / 2, (cudaStream_t)stream);
}

int selfc_haar_rev(const float* z, float* y, int N, int C, int h, int w, void* stream) {
  SELFC_CHECK_ARG(N >= 0 && C >= 1 && h >= 1 && w >= 1, "haar_rev: N=%d C=%d h=%d w=%d", N, C, h, w);
  if (N == 0) return 0;
  SELFC_CHECK_ARG(z && y && ((uintptr_t)y & 7) == 0, "haar_rev: null or misaligned pointer");
  return launch_haar(z, y, true, N, C, h, w, (cudaStream_t)stream);
}

int selfc_gaussian_down(const float* x, const float* k13, float* y, int N, int C, int H, int W, void* stream) {
  SELFC_CHECK_ARG(x && k13 && y, "gaussian_down: null pointer");
  SELFC_CHECK_ARG(N >= 0 && C >= 1 && H >= 16 && W >= 16 && H % 4 == 0 && W % 4 == 0, "gaussian_down: N=%d C=%d H=%d W=%d", N, C, H, W);
  return launch_gaussian_down(x, k13, y, N * C, H, W, (cudaStream_t)stream);
}

int selfc_quantize(const float* x, uint8_t* q_u8, float* q_f32, size_t n, void* stream) {
  SELFC_CHECK_ARG(x || n == 0, "quantize: null input");
  return launch_quantize(x, q_u8, q_f32, n, (cudaStream_t)stream);
}

int selfc_conv3x3(selfc_ctx* ctx, int first_param, int k, const float* x, float* y, int B, int T, int h, int w, void* workspace,
                  size_t workspace_bytes, void* stream) {
  Workspace ws;
  SELFC_TRY(check_run(ctx, B, T, 4 * h, 4 * w, workspace, workspace_bytes, &ws));
  SELFC_CHECK_ARG(x && y && k >= 0 && k < 4, "conv3x3: null pointer or conv index %d outside 0..3", k);
  const DenseW* W = find_dense(ctx, first_param);
  SELFC_CHECK_ARG(W != nullptr, "conv3x3: parameter index %d is not the conv1.weight of a dense block", first_param);
  Dims d{B, T, h, w};
  if (ctx->mode == SELFC_MODE_BF16) return conv3x3_impl<__nv_bfloat16>(ctx, *W, k, x, y, d, (char*)workspace, ws, (cudaStream_t)stream);
  if (ctx->mode == SELFC_MODE_BF16X3) return conv3x3_impl<bfx2>(ctx, *W, k, x, y, d, (char*)workspace, ws, (cudaStream_t)stream);
  return conv3x3_impl<float>(ctx, *W, k, x, y, d, (char*)workspace, ws, (cudaStream_t)stream);
}

int selfc_d2dt(selfc_ctx* ctx, int first_param, const float* x, float* y, int B, int T, int h, int w, void* workspace,
               size_t workspace_bytes, void* stream) {
  Workspace ws;
  SELFC_TRY(check_run(ctx, B, T, 4 * h, 4 * w, workspace, workspace_bytes, &ws));
  SELFC_CHECK_ARG(x && y, "d2dt: null pointer");
  const DenseW* W = find_dense(ctx, first_param);
  SELFC_CHECK_ARG(W != nullptr, "d2dt: parameter index %d is not the conv1.weight of a dense block", first_param);
  Dims d{B, T, h, w};
  if (ctx->mode == SELFC_MODE_BF16) return d2dt_impl<__nv_bfloat16>(ctx, *W, x, y, d, (char*)workspace, ws, (cudaStream_t)stream);
  if (ctx->mode == SELFC_MODE_BF16X3) return d2dt_impl<bfx2>(ctx, *W, x, y, d, (char*)workspace, ws, (cudaStream_t)stream);
  return d2dt_impl<float>(ctx, *W, x, y, d, (char*)workspace, ws, (cudaStream_t)stream);
}

int selfc_global_agg(selfc_ctx* ctx, int first_param, const float* x, float* y, float* wmat_out, int B, int T, int h, int w,
                     void* workspace, size_t workspace_bytes, void* stream) {
  Workspace ws;
  SELFC_TRY(check_run(ctx, B, T, 4 * h, 4 * w, workspace, workspace_bytes, &ws));
  SELFC_CHECK_ARG(x && y, "global_agg: null pointer");
  const int ga_first[6] = {P_GLOBAL1, P_GLOBAL2, P_OTHER + 10, P_OTHER + 28, P_OTHER + 46, P_OTHER + 64};
  const GaW* g = nullptr;
  for (int i = 0; i < 6; ++i)
    if (first_param == ga_first[i]) g = &ctx->ga[i];
  SELFC_CHECK_ARG(g != nullptr, "global_agg: parameter index %d is not the fc.weight of a GlobalAgg", first_param);
  Dims d{B, T, h, w};
  if (ctx->mode == SELFC_MODE_BF16) return ga_impl<__nv_bfloat16>(ctx, *g, x, y, wmat_out, d, (char*)workspace, ws, (cudaStream_t)stream);
  if (ctx->mode == SELFC_MODE_BF16X3) return ga_impl<bfx2>(ctx, *g, x, y, wmat_out, d, (char*)workspace, ws, (cudaStream_t)stream);
  return ga_impl<float>(ctx, *g, x, y, wmat_out, d, (char*)workspace, ws, (cudaStream_t)stream);
}

int selfc_gmm_sample(const float* params, const float* eps, uint64_t seed, uint64_t offset, float* v, int B, int T, int h, int w,
                     void* stream) {
  SELFC_CHECK_ARG(params && v, "gmm_sample: null pointer");
  SELFC_CHECK_ARG(B >= 0 && T >= 1 && h >= 1 && w >= 1, "gmm_sample: bad shape");
  return launch_gmm_sample(params, true, eps, seed, offset, v, true, 0, 0, B, T, h, w, (cudaStream_t)stream);
}

int selfc_gmm_sample_planar(const float* params_planar, const float* eps, uint64_t seed, uint64_t offset, float* z_planar, int B,
                            int T, int h, int w, int form, void* stream) {
  SELFC_CHECK_ARG(params_planar && z_planar, "gmm_sample_planar: null pointer");
  SELFC_CHECK_ARG(B >= 0 && T >= 1 && h >= 1 && w >= 1 && form >= -1 && form <= 1, "gmm_sample_planar: bad shape / form");
  return launch_gmm_sample_planar(params_planar, eps, seed, offset, z_planar, B, T, h, w, (cudaStream_t)stream, form);
}

int selfc_export_eps(float* eps, uint64_t seed, uint64_t offset, int B, int T, int h, int w, void* stream) {
  SELFC_CHECK_ARG(eps, "export_eps: null pointer");
  SELFC_CHECK_ARG(B >= 0 && T >= 1 && h >= 1 && w >= 1, "export_eps: bad shape");
  return launch_export_eps(eps, seed, offset, B, T, (long long)h * w, (cudaStream_t)stream);
}

}  // extern "C"
