// Backward of the training step (models/SelfC_model.py:148-183), templated on the element type of the forward's dense
// buffers: float = FP32 mode (pixel-major buffers, every convolution pass on the fp32-FMA kernel) and bfx2 = BF16X3 mode (slab-planar
// (hi, lo) buffers; forward, input gradients and weight gradients on the tcgen05 kernels: conv_tc3.cu / temporal_tc.cu / wgrad_tc.cu;
// gradients, masks and the optimiser stay fp32).  First building block: the backward pass of one D2DTInput dense block
// (Subnet_constructor.py:115-133), parity-tested against autograd of the oracle.
//
//   forward   x_k = lrelu(conv_k([X, x_1..x_{k-1}])), k = 1..4 (1,3,3);   y = conv5([X, x_1..x_4]) (3,1,1), no activation
//   backward  with gbuf = gradient w.r.t. the dense buffer [X | x1 | x2 | x3 | x4] (zero-initialised):
//     conv5:  dW5 += g_y (x) in5,   gbuf[0:cin5) += conv5^T(g_y)
//     k=4..1: g_k = gbuf[slot_k] * lrelu'(x_k);   dW_k += g_k (x) in_k;   gbuf[0:cin_k) += conv_k^T(g_k)
//     g_X = gbuf[0:cin)
//   conv^T (dgrad) is the SAME implicit-GEMM kernel as the forward conv on the tap-flipped, transposed weights, accumulating into
//   gbuf (conv_simt.cu's EPI_ACCUM / conv_tc3.cu's accumulate epilogue); wgrad is a pixel reduction (wgrad_kernel below / wgrad_tc.cu).
#include <string.h>

#include <stdlib.h>
#include <type_traits>

#include "net_ctx.h"
#include "wgrad_tc.h"

namespace selfc {

constexpr size_t kTrainScratchFloats = 768 * 1024;   // dgrad weights (<= 720 x 256) + weight-gradient scratch (<= 257 x 736), per device

// Training scratch lives in the context (allocated once under its mutex, freed by selfc_ctx_destroy): two contexts, or two
// trainers, on one device never share it.  The context's device is current when these run (check_run).
static float* train_scratch(const selfc_ctx* cctx) {        // dgrad weights + weight-gradient scratch of dense_block_backward
  selfc_ctx* ctx = const_cast<selfc_ctx*>(cctx);
  std::lock_guard<std::mutex> lock(ctx->mu);
  if (!ctx->train_scratch && cudaMalloc(&ctx->train_scratch, kTrainScratchFloats * sizeof(float)) != cudaSuccess) ctx->train_scratch = nullptr;
  return ctx->train_scratch;
}
// planes of the tensor-core weight-gradient kernel for this clip geometry (zeroed when it changes: the padding must read as zero)
static void* train_wg_planes(const selfc_ctx* cctx, const Dims& d, const WgGeom& g, cudaStream_t st) {
  selfc_ctx* ctx = const_cast<selfc_ctx*>(cctx);
  std::lock_guard<std::mutex> lock(ctx->mu);
  const size_t need = wg_plane_bytes(g);
  const bool same = ctx->wg_planes != nullptr && ctx->wg_key[0] == d.B && ctx->wg_key[1] == d.T && ctx->wg_key[2] == d.h && ctx->wg_key[3] == d.w;
  if (same) return ctx->wg_planes;
  if (ctx->wg_bytes < need) {
    if (ctx->wg_planes) cudaFree(ctx->wg_planes);
    ctx->wg_planes = nullptr;
    ctx->wg_bytes = 0;
    if (cudaMalloc(&ctx->wg_planes, need) != cudaSuccess) return nullptr;
    ctx->wg_bytes = need;
  }
  if (cudaMemsetAsync(ctx->wg_planes, 0, need, st) != cudaSuccess) return nullptr;
  ctx->wg_key[0] = d.B; ctx->wg_key[1] = d.T; ctx->wg_key[2] = d.h; ctx->wg_key[3] = d.w;
  return ctx->wg_planes;
}
static float* train_zero_bias(const selfc_ctx* cctx) {      // dgrad has no bias term: 1024 zeros
  selfc_ctx* ctx = const_cast<selfc_ctx*>(cctx);
  std::lock_guard<std::mutex> lock(ctx->mu);
  if (!ctx->train_zero_bias) {
    if (cudaMalloc(&ctx->train_zero_bias, 1024 * sizeof(float)) != cudaSuccess) {
      ctx->train_zero_bias = nullptr;
      return nullptr;
    }
    cudaMemset(ctx->train_zero_bias, 0, 1024 * sizeof(float));
  }
  return ctx->train_zero_bias;
}

// BF16X3 training knobs.  When both the weight and the input gradients run on the tensor cores, the gradient of a dense buffer is kept
// SLAB-PLANAR in fp32 ([pitch / 16][M][16], common.cuh's slab layout with a 4-byte element): the accumulate epilogue of the
// input-gradient launches (thread = pixel, 16 channels = 64 contiguous bytes, consecutive lanes on consecutive rows) and the plane
// builder then move 2 KB runs per warp instead of 32 scattered 64-byte pieces.  Otherwise it is pixel-major [M][pitch] as in FP32 mode.
static bool env_on(const char* name, int& cache) {
  if (cache < 0) {
    const char* e = getenv(name);
    cache = (e && atoi(e) == 0) ? 0 : 1;
  }
  return cache == 1;
}
static bool wgrad_tc_on() { static int c = -1; return env_on("SELFC_WGRAD_TC", c); }
static bool dgrad_tc_on() { static int c = -1; return env_on("SELFC_DGRAD_TC", c); }
static long long grad_slab(const selfc_ctx* ctx, const Dims& d) {
  return ctx->mode == SELFC_MODE_BF16X3 && wgrad_tc_on() && dgrad_tc_on() ? d.M() : 0;
}

// ---- dgrad weights: wd[(tap' * cout4 + n)][c] = w[((taps-1-tap') * cin_buf + c)][n]   (w = forward pack [taps*cin_buf][np]) ----
__global__ void pack_dgrad_kernel(const float* __restrict__ w, float* __restrict__ wd, int taps, int cin_buf, int np, int cout,
                                  int cout4, int npd) {
  const long long total = (long long)taps * cout4 * npd;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int c = (int)(idx % npd);
  const int n = (int)((idx / npd) % cout4);
  const int tp = (int)(idx / ((long long)npd * cout4));
  float v = 0.f;
  if (c < cin_buf && n < cout) v = w[((long long)(taps - 1 - tp) * cin_buf + c) * np + n];
  wd[idx] = v;
}

// g[m][off + n] *= (y[m][off + n] > 0 ? 1 : 0.2)   for n < 32  (LeakyReLU backward through the stored activation)
// E: element type of the forward dense buffer y (float: pixel-major, FP32 mode; bfx2: slab-planar (hi, lo) pairs, BF16X3 mode)
// gslab (optional): the masked gradient also as (hi, lo) slabs [2][M][16 | 16] -- the input of the tensor-core input-gradient launches
template <typename E>
__global__ void lrelu_bwd_kernel(float* __restrict__ g, long long gslabM, const E* __restrict__ y, int pitch, long long slabM, int off, long long M,
                                 bfx2* __restrict__ gslab = nullptr) {
  chain_entry();
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long m = idx >> 3;
  const int c = (int)(idx & 7) * 4;
  if (m >= M) return;
  float* gp = g + dense_off(m, off + c, pitch, gslabM);
  float4 gv = load4(gp);
  const float4 yv = load4(y + dense_off(m, off + c, pitch, slabM));
  gv.x *= yv.x > 0.f ? 1.f : 0.2f;
  gv.y *= yv.y > 0.f ? 1.f : 0.2f;
  gv.z *= yv.z > 0.f ? 1.f : 0.2f;
  gv.w *= yv.w > 0.f ? 1.f : 0.2f;
  store4(gp, gv);
  if (gslab != nullptr) store4(gslab + dense_off(m, c, 32, M), gv);
}

// wgrad: dw[tap][c][n] += sum_m in[m + shift(tap)][c] * g[m][n]  (+ the bias gradient in pseudo-tap `taps`: db[n] += sum_m g[m][n]).
// grid ((taps + 1) * ceil(cout/32), ceil(cin/64), splits); block 256 = 64 input channels x 4 pixel groups.  A CTA reduces its
// share of the pixels through shared-memory tiles of 32 pixels; a thread owns ONE input channel and all 32 output channels of
// the tile (32 accumulators): per pixel one conflict-free LDS of its activation and eight broadcast LDS.128 of the gradient row
// feed 32 FMAs (the first version's 1 x 4 register tile issued 5 LDS per 4 FMAs and ran at a fifth of the FMA rate).  The four
// pixel groups are summed through shared memory, then one atomic per element adds the CTA's partial 64 x 32 block (fp32
// scratch in buffer-channel order).
constexpr int WG_C = 64;
template <typename E>
__global__ void __launch_bounds__(256, 3) wgrad_kernel(const E* __restrict__ in, int in_pitch, long long in_slabM, int cin, const float* __restrict__ g,
                                                    int g_pitch, int g_off, int cout, float* __restrict__ dw, int np, int taps,
                                                    int tap_mode, int BT, int Tn, int h, int w_) {
  __shared__ __align__(16) float As[32][WG_C + 4];     // [pixel][c]
  __shared__ __align__(16) float Gs[32][32];           // [pixel][n]
  __shared__ float Red[WG_C][33];                      // cross-group reduction
  const int tap = blockIdx.x % (taps + 1);
  const int n0 = (blockIdx.x / (taps + 1)) * 32;          // output-channel tile
  const int c0 = blockIdx.y * WG_C;
  const bool bias_pass = tap == taps;
  if (bias_pass && blockIdx.y != 0) return;
  const long long hw = (long long)h * w_;
  const long long M = (long long)BT * hw;
  const int tid = threadIdx.x;
  const int tc = tid & (WG_C - 1), pg = tid >> 6;
  int dy = 0, dx = 0, dt = 0;
  if (!bias_pass) {
    if (tap_mode == TAP_SPATIAL) { dy = tap / 3 - 1; dx = tap % 3 - 1; }
    else if (tap_mode == TAP_TEMPORAL) dt = tap - 1;
  }
  float acc[32];
#pragma unroll
  for (int j = 0; j < 32; ++j) acc[j] = 0.f;
  const long long per = (M + gridDim.z - 1) / gridDim.z;
  const long long m_begin = (long long)blockIdx.z * per;
  const long long m_end = m_begin + per < M ? m_begin + per : M;
  // global -> registers for the tile starting at mb (the loads of tile i+1 are in flight while tile i is being reduced)
  float4 rg, ra[2];
  auto fetch = [&](long long mb) {
    {   // gradient rows: thread (p = tid/8, q = tid%8) loads 4 output channels
      const int p = tid >> 3, q4 = (tid & 7) * 4;
      const long long m = mb + p;
      rg = make_float4(0.f, 0.f, 0.f, 0.f);
      if (m < m_end && n0 + q4 < cout) {
        const float* gp = g + m * g_pitch + g_off + n0 + q4;
        rg.x = gp[0];
        if (n0 + q4 + 1 < cout) rg.y = gp[1];
        if (n0 + q4 + 2 < cout) rg.z = gp[2];
        if (n0 + q4 + 3 < cout) rg.w = gp[3];
      }
    }
#pragma unroll
    for (int i = 0; i < 2; ++i) {   // (shifted) activations: 32 pixels x 16 float4
      const int idx = tid + i * 256;
      const int p = idx >> 4, q4 = (idx & 15) * 4;
      const long long m = mb + p;
      ra[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (m < m_end) {
        if (bias_pass) {
          ra[i] = make_float4(1.f, 1.f, 1.f, 1.f);
        } else if (c0 + q4 < cin) {
          const long long n = m / hw, pix = m - n * hw;
          const int y = (int)(pix / w_), x = (int)(pix - (long long)y * w_);
          const int t = (int)(n % Tn);
          const bool ok = (unsigned)(y + dy) < (unsigned)h && (unsigned)(x + dx) < (unsigned)w_ && (unsigned)(t + dt) < (unsigned)Tn;
          if (ok) ra[i] = load4(in + dense_off(m + (long long)dt * hw + dy * w_ + dx, c0 + q4, in_pitch, in_slabM));
        }
      }
    }
  };
  if (m_begin < m_end) fetch(m_begin);
  for (long long mb = m_begin; mb < m_end; mb += 32) {
    *reinterpret_cast<float4*>(&Gs[tid >> 3][(tid & 7) * 4]) = rg;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int idx = tid + i * 256;
      *reinterpret_cast<float4*>(&As[idx >> 4][(idx & 15) * 4]) = ra[i];
    }
    __syncthreads();
    if (mb + 32 < m_end) fetch(mb + 32);
#pragma unroll 2
    for (int pp = pg; pp < 32; pp += 4) {
      const float a = As[pp][tc];
#pragma unroll
      for (int j4 = 0; j4 < 8; ++j4) {
        const float4 gv = *reinterpret_cast<const float4*>(&Gs[pp][j4 * 4]);
        acc[4 * j4 + 0] = fmaf(a, gv.x, acc[4 * j4 + 0]);
        acc[4 * j4 + 1] = fmaf(a, gv.y, acc[4 * j4 + 1]);
        acc[4 * j4 + 2] = fmaf(a, gv.z, acc[4 * j4 + 2]);
        acc[4 * j4 + 3] = fmaf(a, gv.w, acc[4 * j4 + 3]);
      }
    }
    __syncthreads();
  }
  // sum the four pixel groups into group 0
  for (int gsrc = 1; gsrc < 4; ++gsrc) {
    if (pg == gsrc) {
#pragma unroll
      for (int j = 0; j < 32; ++j) Red[tc][j] = acc[j];
    }
    __syncthreads();
    if (pg == 0) {
#pragma unroll
      for (int j = 0; j < 32; ++j) acc[j] += Red[tc][j];
    }
    __syncthreads();
  }
  if (pg != 0) return;
  if (bias_pass) {
    if (tc == 0)
      for (int j = 0; j < 32; ++j)
        if (n0 + j < cout) atomicAdd(dw + (long long)taps * cin * np + n0 + j, acc[j]);
  } else if (c0 + tc < cin) {
    for (int j = 0; j < 32; ++j)
      if (n0 + j < cout) atomicAdd(dw + ((long long)tap * cin + c0 + tc) * np + n0 + j, acc[j]);
  }
}

// pixel splits of a wgrad launch: the grid should fill the resident slots (3 CTAs per SM) once or twice, not 1.1 times
static int wgrad_splits(long long M, int base_ctas) {
  int dev = 0, nsm = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev);
  const int slots = 3 * nsm;
  int waves = M >= 400000 ? 4 : (M >= 100000 ? 2 : 1);
  int splits = waves * slots / (base_ctas > 0 ? base_ctas : 1);
  const long long max_splits = M / 256 > 0 ? M / 256 : 1;      // at least 8 tiles of 32 pixels per CTA
  if (splits > max_splits) splits = (int)max_splits;
  if (splits > 1024) splits = 1024;
  if (splits < 1) splits = 1;
  return splits;
}

// scratch [taps][cin_buf][np] (+ [np] bias) in buffer-channel order -> reference layouts dW [cout][cin_ref][taps], db [cout] (accumulating)
__global__ void wgrad_unpack_kernel(const float* __restrict__ dw, float* __restrict__ gw, float* __restrict__ gb, int cout, int cin_ref,
                                    int taps, int cin_buf, int xreal, int xpad, int np) {
  chain_entry();
  const long long total = (long long)cout * cin_ref * taps;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx < cout) gb[idx] += dw[(long long)taps * cin_buf * np + idx];
  if (idx >= total) return;
  const int tap = (int)(idx % taps);
  const int cref = (int)((idx / taps) % cin_ref);
  const int n = (int)(idx / ((long long)taps * cin_ref));
  const int c = cref < xreal ? cref : cref - xreal + xpad;
  gw[idx] += dw[((long long)tap * cin_buf + c) * np + n];
}

// NCHW [N,C,h,w] -> pixel-major [M][pitch] (zero-filled to cpad) and back, fp32 (thin wrappers over layout.cu)
static int to_dense(const float* x, float* dst, int pitch, int C, int cpad, const Dims& d, cudaStream_t st) {
  return launch_nchw_to_dense<float>(x, dst, pitch, 0, 0, C, cpad, d.M(), d.hw(), st);
}

// ---- backward of one dense block ----------------------------------------------------------------------------------
// buf: the block's forward dense buffer [M][pitch] (X, x1..x4 as left by the forward convs); gy: [M][gy_pitch] gradient of
// the block output (channels [0,cout), zero padded to a multiple of 4); gbuf: [M][pitch] scratch, receives the gradient
// of the dense buffer (g_X = its first W.xpad channels); scratch: >= wgrad_scratch_floats(W) floats;
// gparams[10]: conv1.weight, conv1.bias, ..., conv5.bias gradients in reference layout, ACCUMULATED into (may be null: skip wgrad).
size_t dense_bwd_scratch_floats(const DenseW& W) {
  const int cin5 = W.xpad + 4 * kGrowth;
  size_t wd = 0, dw = 0;
  for (int k = 0; k < 5; ++k) {
    const int taps = k < 4 ? 9 : 3;
    const int cin = W.xpad + kGrowth * k;
    const int cout = k < 4 ? kGrowth : W.cout;
    const int cout4 = (cout + 3) & ~3;
    const int npd = (cin + 31) & ~31;
    const size_t a = (size_t)taps * cout4 * npd, b = (size_t)(taps * cin + 1) * W.np[k];
    wd = a > wd ? a : wd;
    dw = b > dw ? b : dw;
  }
  (void)cin5;
  return wd + dw + 256;
}

// ---- tensor-core input gradients (BF16X3 mode) -------------------------------------------------------------------------
// conv5: wd5[c][n][dt] = Wf5[((2 - dt) * cin_buf + c) * np + n] (n < cout, else 0): the flipped / transposed temporal weights in the
// reference layout [cout' = buffer channel][cin' = nb][3] that pack_temporal_weights takes
struct Dgrad5RefJob {
  const float* wf;
  float* wd5;
  int cin_buf, np, cout, nb;
};
__device__ __forceinline__ void dgrad5_ref_body(const Dgrad5RefJob& j, int idx) {
  const int total = j.cin_buf * j.nb * 3;
  if (idx >= total) return;
  const int dt = idx % 3, n = (idx / 3) % j.nb, c = idx / (3 * j.nb);
  j.wd5[idx] = n < j.cout ? j.wf[((size_t)(2 - dt) * j.cin_buf + c) * j.np + n] : 0.f;
}
__global__ void dgrad5_ref_kernel(const Dgrad5RefJob j) { dgrad5_ref_body(j, blockIdx.x * blockDim.x + threadIdx.x); }
// every recorded job in one launch (pack_batch.h)
__global__ void dgrad5_ref_multi_kernel(const Dgrad5RefJob* __restrict__ jobs, const int* __restrict__ first, int njobs) {
  const int ji = pack_find_job(first, njobs, blockIdx.x);
  const Dgrad5RefJob j = jobs[ji];
  dgrad5_ref_body(j, (blockIdx.x - __ldg(first + ji)) * blockDim.x + threadIdx.x);
}
int flush_pack_dgrad5_ref(JobTable& t, cudaStream_t st) {
  if (t.njobs() <= 0) return 0;
  const void* jobs = nullptr;
  const int* first = nullptr;
  SELFC_CUDA(t.sync(st, &jobs, &first));
  dgrad5_ref_multi_kernel<<<t.first.back(), 256, 0, st>>>(static_cast<const Dgrad5RefJob*>(jobs), first, t.njobs());
  SELFC_LAUNCH_CHECK("dgrad5_ref_multi_kernel");
  return 0;
}
// fp32 pixel-major [M][spitch] columns [0, ncol) -> (hi, lo) slabs [nb / 16][M], columns >= ncol zero
__global__ void cols_to_slab_kernel(bfx2* __restrict__ dst, int nb, const float* __restrict__ src, int spitch, int ncol, long long M) {
  chain_entry();
  const int per = nb / 4;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= M * per) return;
  const long long m = idx / per;
  const int c = (int)(idx - m * per) * 4;
  float v[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) v[e] = c + e < ncol ? src[m * spitch + c + e] : 0.f;
  store4(dst + dense_off(m, c, nb, M), make_float4(v[0], v[1], v[2], v[3]));
}

static int ensure_dgrad_images(const selfc_ctx* cctx, const DenseW& cW, float* zero_bias, cudaStream_t st) {
  DenseW& W = const_cast<DenseW&>(cW);
  if (W.dg_valid) return 0;
  selfc_ctx* ctx = const_cast<selfc_ctx*>(cctx);
  std::lock_guard<std::mutex> lock(ctx->mu);
  // slot s of the dense buffer (s = 0: X, s >= 1: x_s) receives conv_{s+1..4}^T: 4 - s contributing convs
  const float* wf[4] = {W.w[0], W.w[1], W.w[2], W.w[3]};
  const int cins[4] = {W.xpad, W.xpad + kGrowth, W.xpad + 2 * kGrowth, W.xpad + 3 * kGrowth};
  for (int sl = 0; sl < 4; ++sl) {
    const int c0 = sl == 0 ? 0 : W.xpad + kGrowth * (sl - 1), ncover = sl == 0 ? W.xpad : kGrowth, nconv = 4 - sl;
    if (W.dg_img[sl] == nullptr) SELFC_CUDA(cudaMalloc(&W.dg_img[sl], tc3_dgrad_slot_image_bytes(nconv) * cdiv(ncover, 32)));
    SELFC_TRY(pack_tc3_dgrad_slot_images(wf, cins, W.dg_img[sl], c0, ncover, nconv, st));
  }
  const int cin5 = W.xpad + 4 * kGrowth, nb = (W.cout + 15) & ~15;
  // (one scratch per block: with the packs batched (pack_batch.h) every block's flipped weights exist at the same time)
  if (W.dg5_tmp == nullptr) SELFC_CUDA(cudaMalloc(&W.dg5_tmp, (size_t)192 * 64 * 3 * sizeof(float)));
  const Dgrad5RefJob rj{W.w[4], W.dg5_tmp, cin5, W.np[4], W.cout, nb};
  if (PackBatch* pb = pack_batch_current()) {
    pb->ref5[pb->point].add(rj, (int)cdiv(cin5 * nb * 3, 256));
  } else {
    dgrad5_ref_kernel<<<cdiv(cin5 * nb * 3, 256), 256, 0, st>>>(rj);
    SELFC_LAUNCH_CHECK("dgrad5_ref_kernel");
  }
  W.dg5_c0[0] = 0; W.dg5_n[0] = cin5 > 96 ? 96 : cin5;
  W.dg5_c0[1] = W.dg5_n[0]; W.dg5_n[1] = cin5 - W.dg5_n[0];
  for (int gI = 0; gI < 2; ++gI)
    if (W.dg5_n[gI] > 0)
      SELFC_TRY(pack_temporal_weights(W.dg5[gI], W.dg5_tmp + (size_t)W.dg5_c0[gI] * nb * 3, zero_bias, W.dg5_n[gI], nb, 3, nb, nb, nb, st, true));
  W.dg_valid = true;
  return 0;
}
static bfx2* train_gslab(const selfc_ctx* cctx, long long M) {
  selfc_ctx* ctx = const_cast<selfc_ctx*>(cctx);
  std::lock_guard<std::mutex> lock(ctx->mu);
  // 64 channels for a dense-block conv's output gradient; 64 + 128 + 256 + 720 + 256 + 128 for the GMM head's tensors (head_sampler_backward)
  const size_t need = (size_t)M * 1600 * sizeof(bfx2) + 4096;
  if (ctx->dg_gslab_bytes < need) {
    if (ctx->dg_gslab) cudaFree(ctx->dg_gslab);
    ctx->dg_gslab = nullptr;
    ctx->dg_gslab_bytes = 0;
    if (cudaMalloc(&ctx->dg_gslab, need) != cudaSuccess) return nullptr;
    ctx->dg_gslab_bytes = need;
  }
  return reinterpret_cast<bfx2*>(ctx->dg_gslab);
}

template <typename E>
int dense_block_backward(const selfc_ctx* ctx, const DenseW& W, const E* buf, int pitch, const float* gy, int gy_pitch, float* gbuf,
                         float* scratch, float* const* gparams, const Dims& d, cudaStream_t st) {
  const long long M = d.M();
  const long long slabM = dense_slab(ctx, d);         // layout of the forward buffer; gradient buffers are fp32 pixel-major
  if (M == 0) return 0;
  float* zero_bias = train_zero_bias(ctx);     // dgrad has no bias term
  SELFC_CHECK_ARG(zero_bias != nullptr, "out of device memory (training scratch)");
  // BF16X3 mode: weight gradients on the tensor cores (wgrad_tc.cu) from transposed, zero-padded planes of the block's activations,
  // built once for all five convolutions.  SELFC_WGRAD_TC=0: the fp32-FMA pixel reduction below (A/B, parity).
  bool wg_tc = false;
  WgGeom geom{};
  void* planes = nullptr;
  if constexpr (std::is_same<E, bfx2>::value) {
    if (wgrad_tc_on() && gparams != nullptr) {
      geom = wg_geometry(d);
      planes = train_wg_planes(ctx, d, geom, st);
      SELFC_CHECK_ARG(planes != nullptr, "out of device memory (weight-gradient planes, %zu bytes)", wg_plane_bytes(geom));
      SELFC_TRY(launch_wg_planes_act(buf, pitch, d, geom, planes, st));
      wg_tc = true;
    }
  }
  // ... and the input gradients: the tcgen05 forward kernels on the flipped / transposed (hi, lo) weight images, the output gradient
  // of each conv converted to (hi, lo) slabs on the way.  SELFC_DGRAD_TC=0: the fp32-FMA kernel.
  bool dg_tc = false;
  bfx2* gslab = nullptr;
  if constexpr (std::is_same<E, bfx2>::value) {
    if (dgrad_tc_on()) {
      SELFC_TRY(ensure_dgrad_images(ctx, W, zero_bias, st));
      gslab = train_gslab(ctx, M);
      SELFC_CHECK_ARG(gslab != nullptr, "out of device memory (gradient slabs)");
      dg_tc = true;
    }
  }
  const long long gslabM = grad_slab(ctx, d);         // layout of gbuf (see grad_slab); implies wg_tc || gparams == nullptr, and dg_tc
  // (tensor-core input gradients: conv5's launches STORE all xpad + 128 buffer channels of every pixel first, nothing to clear)
  if (!dg_tc) SELFC_CUDA(cudaMemsetAsync(gbuf, 0, (size_t)M * pitch * sizeof(float), st));
  for (int k = 4; k >= 0; --k) {
    const int taps = k < 4 ? 9 : 3;
    const int tap_mode = k < 4 ? TAP_SPATIAL : TAP_TEMPORAL;
    const int cin = W.xpad + kGrowth * k;                  // buffer channels consumed by conv_{k+1}
    const int cout = k < 4 ? kGrowth : W.cout;
    const int cout4 = (cout + 3) & ~3;
    const int npd = (cin + 31) & ~31;
    const int slot = W.xpad + kGrowth * k;                 // where x_{k+1} lives (k < 4)
    const float* g = k < 4 ? gbuf : gy;
    const int g_pitch = k < 4 ? pitch : gy_pitch;
    const int g_off = k < 4 ? slot : 0;
    if (k < 4) {
      // (the masked gradients of conv4, conv3, ... are kept side by side as (hi, lo) slabs: the K operand of the input-gradient launches)
      SELFC_CUDA(launch_chain(lrelu_bwd_kernel<E>, dim3((unsigned)cdiv(M * 8, 256)), 256, 0, st, gbuf, gslabM, buf, pitch, slabM, slot, M,
                              dg_tc ? gslab + (size_t)(2 * (3 - k)) * M * 16 : static_cast<bfx2*>(nullptr)));
      SELFC_LAUNCH_CHECK("lrelu_bwd_kernel");
    }
    float* wd = scratch;
    float* dw = scratch + (size_t)taps * cout4 * npd;
    // weight / bias gradients
    if (gparams != nullptr && gparams[2 * k] != nullptr) {
      const size_t dw_floats = (size_t)(taps * cin + 1) * W.np[k];
      if (wg_tc) {
        // (the plane builder clears the split-K scratch on its way: mask -> planes -> weight gradient -> unpack -> input gradient stays one
        // chain of kernels, launch_chain)
        const int nb = k < 4 ? kGrowth : (cout + 15) & ~15;
        SELFC_TRY(launch_wg_planes_grad(g, g_pitch, g_off, k < 4 ? gslabM : 0, cout, nb, k == 4, d, geom, planes, st, dw, (long long)dw_floats));
        SELFC_TRY(launch_wgrad_tc(planes, geom, cin, cout, nb, k == 4 ? WG_TEMPORAL : WG_SPATIAL, dw, W.np[k], st));
      } else {
        SELFC_CUDA(cudaMemsetAsync(dw, 0, dw_floats * sizeof(float), st));
        SELFC_CHECK_ARG(gslabM == 0, "the fp32-FMA weight gradient reads a pixel-major gradient buffer");
        const int splits = wgrad_splits(M, (taps + 1) * cdiv(cout, 32) * cdiv(cin, WG_C));
        dim3 grid((taps + 1) * cdiv(cout, 32), cdiv(cin, WG_C), splits);
        wgrad_kernel<E><<<grid, 256, 0, st>>>(buf, pitch, slabM, cin, g, g_pitch, g_off, cout, dw, W.np[k], taps, tap_mode, d.B * d.T, d.T, d.h, d.w);
        SELFC_LAUNCH_CHECK("wgrad_kernel");
      }
      const int cin_ref = W.cin + kGrowth * k;
      const long long total = (long long)cout * cin_ref * taps;
      SELFC_CUDA(launch_chain(wgrad_unpack_kernel, dim3((unsigned)cdiv(total > cout ? total : cout, 256)), 256, 0, st, (const float*)dw, gparams[2 * k],
                              gparams[2 * k + 1], cout, cin_ref, taps, cin, W.cin, W.xpad, W.np[k]));
      SELFC_LAUNCH_CHECK("wgrad_unpack_kernel");
    }
    // input gradient: the forward kernel on the flipped / transposed weights, accumulated into gbuf[0:cin)
    if (dg_tc && k == 4) {
      // conv5: gy -> (hi, lo) slabs, then the temporal kernel per column group, STORING into the still-zero gradient buffer
      const int nb = (cout + 15) & ~15;
      SELFC_CUDA(launch_chain(cols_to_slab_kernel, dim3((unsigned)cdiv(M * (nb / 4), 256)), 256, 0, st, gslab, nb, gy, gy_pitch, cout, M));
      SELFC_LAUNCH_CHECK("cols_to_slab_kernel");
      for (int gI = 0; gI < 2; ++gI) {
        if (W.dg5_n[gI] <= 0) continue;
        TcTempArgs t;
        t.in = reinterpret_cast<const __nv_bfloat16*>(gslab); t.in_pitch = nb; t.B = d.B; t.T = d.T; t.hw = (int)d.hw(); t.in_slabM = M;
        t.epi = EPI_STORE; t.act = 0;
        t.outF = gbuf; t.outF_pitch = pitch; t.outF_off = W.dg5_c0[gI]; t.outF_slabM = gslabM;
        SELFC_TRY(launch_temporal_tc(W.dg5[gI], t, st));
      }
      continue;
    }
    if (dg_tc) {
      // slot by slot: with conv_{k+1}'s masked output gradient in place, the buffer slot it reads LAST -- x_k, or X for k = 0 -- has all
      // its contributions available: conv5's (stored above) + one launch with the gradients of conv4..conv_{k+1} concatenated in K
      // (K = 32 (4 - k)), accumulated ONCE.  (Conv by conv, a channel was read-modified-written once per later conv.)
      const int nconv = 4 - k;
      const int c0 = k == 0 ? 0 : W.xpad + kGrowth * (k - 1), ncover = k == 0 ? W.xpad : kGrowth;
      TcConvW tw;
      tw.img = W.dg_img[k];                   // presence only: the (hi, lo) form reads img_x2
      tw.img_x2 = W.dg_img[k];
      tw.img_bytes = tc3_dgrad_slot_image_bytes(nconv) / 2;
      tw.cin_buf = 32 * nconv;
      tw.bias = zero_bias;
      TcAccum acc;
      acc.out = gbuf; acc.pitch = pitch; acc.off = c0; acc.n = ncover; acc.ngroups = cdiv(ncover, 32); acc.slabM = gslabM;
      SELFC_TRY(launch_conv3x3_tc(tw, reinterpret_cast<__nv_bfloat16*>(gslab), M, 32 * nconv, 0, d.B * d.T, d.h, d.w, st, nullptr, nullptr, true, &acc));
      continue;
    }
    const long long wtot = (long long)taps * cout4 * npd;
    pack_dgrad_kernel<<<cdiv(wtot, 256), 256, 0, st>>>(W.w[k], wd, taps, cin, W.np[k], cout, cout4, npd);
    SELFC_LAUNCH_CHECK("pack_dgrad_kernel");
    ConvArgs<float> a;
    a.in = g + g_off; a.in_pitch = g_pitch; a.cin = cout4;
    a.w = wd; a.bias = zero_bias; a.np = npd; a.cout = cin;
    a.taps = taps; a.tap_mode = tap_mode;
    a.BT = d.B * d.T; a.Tn = d.T; a.h = d.h; a.w_ = d.w;
    a.epi = EPI_ACCUM; a.outF = gbuf; a.outF_pitch = pitch; a.outF_off = 0;
    SELFC_TRY(launch_conv_simt<float>(a, st));
  }
  return 0;
}


// ====================================================================================================================
// InvBlockExp backward (SelfC_GMM_arch_inv.py:21-33), both directions.  State and its gradient are planar quads
// [13][M][4] (common.cuh); the block's forward is RECOMPUTED from its saved input state (invblock_fwd), which leaves the
// three dense buffers and the log-scale s in the workspace.
//   forward dir:  y1 = x1 + F(x2);  s = 2*sigmoid(H(y1))-1;  y2 = x2*e^s + G(y1)
//   reverse dir:  s = 2*sigmoid(H(x1))-1;  y2 = (x2 - G(x1))*e^-s;  y1 = x1 - F(y2)
// ====================================================================================================================
// quads [q0, q0 + nq) of the planar state -> channels [off, off + 4 nq) of a dense buffer; channels up to off + cpad are zero-filled
// (the 16-channel X slot of the 3-input blocks in the tensor-core modes)
template <typename E>
__global__ void quads_to_dense_kernel(const float* __restrict__ z, int q0, int nq, E* __restrict__ dst, int pitch, long long slabM, int off,
                                      int cpad, long long M) {
  const int per = cpad / 4 > nq ? cpad / 4 : nq;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= M * per) return;
  const long long m = idx / per;
  const int j = (int)(idx - m * per);
  const float4 v = j < nq ? load4(z + quad_off((size_t)M, q0 + j, (size_t)m)) : make_float4(0.f, 0.f, 0.f, 0.f);
  store4(dst + dense_off(m, off + 4 * j, pitch, slabM), v);
}

// forward dir, step 1: gyG = gy2, gyH = gy2*x2*e^s*(1-s^2)/2, gz[x2 part] = gy2*e^s
__global__ void cpl_fwd_pre_kernel(float* __restrict__ gz, const float* __restrict__ zin, const float* __restrict__ sbuf,
                                   float* __restrict__ gyG, float* __restrict__ gyH, long long M) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= M * kSQuads) return;
  const long long m = idx / kSQuads;
  const int q = (int)(idx - m * kSQuads);
  const float4 g = load4(gz + quad_off((size_t)M, 1 + q, (size_t)m));
  const float4 x = load4(zin + quad_off((size_t)M, 1 + q, (size_t)m));
  const float4 s = load4(sbuf + quad_off((size_t)M, q, (size_t)m));
  const float gg[4] = {g.x, g.y, g.z, g.w}, xx[4] = {x.x, x.y, x.z, x.w}, ss[4] = {s.x, s.y, s.z, s.w};
  float oG[4], oH[4], oX[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const float ex = expf(ss[e]);
    oG[e] = gg[e];
    oH[e] = gg[e] * xx[e] * ex * (1.0f - ss[e] * ss[e]) * 0.5f;
    oX[e] = gg[e] * ex;
  }
  store4(gyG + m * kHF + 4 * q, make_float4(oG[0], oG[1], oG[2], oG[3]));
  store4(gyH + m * kHF + 4 * q, make_float4(oH[0], oH[1], oH[2], oH[3]));
  store4(gz + quad_off((size_t)M, 1 + q, (size_t)m), make_float4(oX[0], oX[1], oX[2], oX[3]));
}

// acc[m][0:4] (+)= src[m][0:4]  (first channels of a dense-buffer gradient); init: overwrite instead of add
__global__ void take4_kernel(float* __restrict__ acc, const float* __restrict__ src, int pitch, long long sslabM, int init, float sign, long long M) {
  const long long m = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= M) return;
  const float4 v = load4(src + dense_off(m, 0, pitch, sslabM));
  float4 a = init ? make_float4(0.f, 0.f, 0.f, 0.f) : load4(acc + m * 4);
  a.x += sign * v.x; a.y += sign * v.y; a.z += sign * v.z; a.w = 0.f;
  store4(acc + m * 4, a);
}

// gz quad 0 += acc (xyz);  optionally also out4[m] = sign * gz quad 0 (the 4-channel gradient fed to F's backward)
__global__ void cpl_quad0_kernel(float* __restrict__ gz, const float* __restrict__ acc, float* __restrict__ out4, float sign, long long M) {
  const long long m = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= M) return;
  float4 g = load4(gz + quad_off((size_t)M, 0, (size_t)m));
  if (acc) {
    const float4 a = load4(acc + m * 4);
    g.x += a.x; g.y += a.y; g.z += a.z;
  }
  g.w = 0.f;
  store4(gz + quad_off((size_t)M, 0, (size_t)m), g);
  if (out4) store4(out4 + m * 4, make_float4(sign * g.x, sign * g.y, sign * g.z, 0.f));
}

// gz[x2 part] += gX[m][0:48]   (forward dir, after F's backward)
__global__ void cpl_add_hf_kernel(float* __restrict__ gz, const float* __restrict__ gX, int pitch, long long xslabM, long long M) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= M * kSQuads) return;
  const long long m = idx / kSQuads;
  const int q = (int)(idx - m * kSQuads);
  float4 g = load4(gz + quad_off((size_t)M, 1 + q, (size_t)m));
  const float4 a = load4(gX + dense_off(m, 4 * q, pitch, xslabM));
  g.x += a.x; g.y += a.y; g.z += a.z; g.w += a.w;
  store4(gz + quad_off((size_t)M, 1 + q, (size_t)m), g);
}

// reverse dir: t = gy2 + gXF;  gz[x2 part] = t*e^-s;  gyG = -t*e^-s;  gyH = -t*y2*(1-s^2)/2
__global__ void cpl_rev_pre_kernel(float* __restrict__ gz, const float* __restrict__ gXF, int pitchF, long long fslabM, const float* __restrict__ zout,
                                   const float* __restrict__ sbuf, float* __restrict__ gyG, float* __restrict__ gyH, long long M) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= M * kSQuads) return;
  const long long m = idx / kSQuads;
  const int q = (int)(idx - m * kSQuads);
  const float4 g = load4(gz + quad_off((size_t)M, 1 + q, (size_t)m));
  const float4 f = load4(gXF + dense_off(m, 4 * q, pitchF, fslabM));
  const float4 y = load4(zout + quad_off((size_t)M, 1 + q, (size_t)m));
  const float4 s = load4(sbuf + quad_off((size_t)M, q, (size_t)m));
  const float tt[4] = {g.x + f.x, g.y + f.y, g.z + f.z, g.w + f.w}, yy[4] = {y.x, y.y, y.z, y.w}, ss[4] = {s.x, s.y, s.z, s.w};
  float oG[4], oH[4], oX[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const float ex = expf(-ss[e]);
    oX[e] = tt[e] * ex;
    oG[e] = -tt[e] * ex;
    oH[e] = -tt[e] * yy[e] * (1.0f - ss[e] * ss[e]) * 0.5f;
  }
  store4(gyG + m * kHF + 4 * q, make_float4(oG[0], oG[1], oG[2], oG[3]));
  store4(gyH + m * kHF + 4 * q, make_float4(oH[0], oH[1], oH[2], oH[3]));
  store4(gz + quad_off((size_t)M, 1 + q, (size_t)m), make_float4(oX[0], oX[1], oX[2], oX[3]));
}


// zin: the block's saved input state (planar); gz: gradient w.r.t. the block's OUTPUT state on entry, w.r.t. its INPUT state
// on return; gparams: 30 gradients (F, G, H x conv1..5 weight/bias), accumulated into.  Scratch inside the workspace: the STP
// regions (params, h1, h2), which are idle while a coupling block runs.
// kept (optional): the block's buffers as its forward left them + zout = the block's output state; then nothing is recomputed
template <typename E>
int invblock_backward(const selfc_ctx* ctx, int blk, bool rev, const float* zin, float* gz, float* const* gparams, char* wsp,
                      const Workspace& ws, const Dims& d, cudaStream_t st, const BlockBufsV* kept = nullptr, const float* zout = nullptr) {
  const long long M = d.M();
  const long long slabM = dense_slab(ctx, d);
  const int xp3 = ctx->xpad3;
  float* z = reinterpret_cast<float*>(wsp + ws.z);
  float* sbuf = kept ? kept->s : reinterpret_cast<float*>(wsp + ws.sbuf);
  E* fbuf = kept ? static_cast<E*>(kept->f) : reinterpret_cast<E*>(wsp + ws.fbuf);
  E* gbuf = kept ? static_cast<E*>(kept->g) : reinterpret_cast<E*>(wsp + ws.gbuf);
  E* hbuf = kept ? static_cast<E*>(kept->h) : reinterpret_cast<E*>(wsp + ws.hbuf);
  const float* zo = kept ? zout : z;                  // the block's output state (the reverse direction's coupling gradient reads y2)
  float* gdense = reinterpret_cast<float*>(wsp + ws.params);                 // [M][<=192] gradient of a dense buffer
  float* gyG = reinterpret_cast<float*>(wsp + ws.h2);                        // [M][48]
  float* gyH = gyG + (size_t)M * kHF;                                        // [M][48]
  float* gyF = reinterpret_cast<float*>(wsp + ws.h1);                        // [M][4]
  float* acc = gyF + (size_t)M * 4;                                          // [M][4]
  float* scratch = train_scratch(ctx);
  SELFC_CHECK_ARG(scratch != nullptr, "out of device memory (training scratch)");
  const DenseW& F = ctx->inv[blk][0];
  const DenseW& G = ctx->inv[blk][1];
  const DenseW& H = ctx->inv[blk][2];
  float* const* gF = gparams ? gparams : nullptr;
  float* const* gG = gparams ? gparams + 10 : nullptr;
  float* const* gH = gparams ? gparams + 20 : nullptr;
  const int nb = cdiv(M, 256), nbq = cdiv(M * kSQuads, 256), nbx = cdiv(M * (xp3 / 4), 256);
  const long long gsl = grad_slab(ctx, d);           // layout of the dense-buffer gradients dense_block_backward leaves in gdense
  // recompute the block's forward from its input state (unless its buffers were kept)
  if (kept == nullptr) {
  SELFC_CUDA(cudaMemcpyAsync(z, zin, (size_t)M * kZQuads * 16, cudaMemcpyDeviceToDevice, st));
  if (!rev) {
    quads_to_dense_kernel<E><<<nbq, 256, 0, st>>>(zin, 1, kSQuads, fbuf, ws.fpitch, slabM, 0, 0, M);
    SELFC_LAUNCH_CHECK("quads_to_dense_kernel");
  } else {
    quads_to_dense_kernel<E><<<nbx, 256, 0, st>>>(zin, 0, 1, gbuf, ws.gpitch, slabM, 0, xp3, M);
    quads_to_dense_kernel<E><<<nbx, 256, 0, st>>>(zin, 0, 1, hbuf, ws.gpitch, slabM, 0, xp3, M);
    SELFC_LAUNCH_CHECK("quads_to_dense_kernel");
  }
  SELFC_TRY(invblock_fwd<E>(ctx, blk, rev, wsp, ws, d, st));
  // the block's last epilogue has already put the NEXT block's input into an X slot (y2 into F's, or y1 into G's and H's):
  // restore this block's own inputs, which the weight gradients need
  if (!rev) {
    quads_to_dense_kernel<E><<<nbq, 256, 0, st>>>(zin, 1, kSQuads, fbuf, ws.fpitch, slabM, 0, 0, M);
  } else {
    quads_to_dense_kernel<E><<<nbx, 256, 0, st>>>(zin, 0, 1, gbuf, ws.gpitch, slabM, 0, xp3, M);
    quads_to_dense_kernel<E><<<nbx, 256, 0, st>>>(zin, 0, 1, hbuf, ws.gpitch, slabM, 0, xp3, M);
  }
  SELFC_LAUNCH_CHECK("quads_to_dense_kernel");
  }
  if (!rev) {
    cpl_fwd_pre_kernel<<<nbq, 256, 0, st>>>(gz, zin, sbuf, gyG, gyH, M);
    SELFC_LAUNCH_CHECK("cpl_fwd_pre_kernel");
    SELFC_TRY(dense_block_backward<E>(ctx, G, gbuf, ws.gpitch, gyG, kHF, gdense, scratch, gG, d, st));
    take4_kernel<<<nb, 256, 0, st>>>(acc, gdense, ws.gpitch, gsl, 1, 1.0f, M);
    SELFC_TRY(dense_block_backward<E>(ctx, H, hbuf, ws.gpitch, gyH, kHF, gdense, scratch, gH, d, st));
    take4_kernel<<<nb, 256, 0, st>>>(acc, gdense, ws.gpitch, gsl, 0, 1.0f, M);
    cpl_quad0_kernel<<<nb, 256, 0, st>>>(gz, acc, gyF, 1.0f, M);             // gx1 = gy1 + G^T + H^T; F's output gradient = the same
    SELFC_LAUNCH_CHECK("cpl_quad0_kernel");
    SELFC_TRY(dense_block_backward<E>(ctx, F, fbuf, ws.fpitch, gyF, 4, gdense, scratch, gF, d, st));
    cpl_add_hf_kernel<<<nbq, 256, 0, st>>>(gz, gdense, ws.fpitch, gsl, M);
    SELFC_LAUNCH_CHECK("cpl_add_hf_kernel");
  } else {
    cpl_quad0_kernel<<<nb, 256, 0, st>>>(gz, nullptr, gyF, -1.0f, M);        // y1 = x1 - F(y2): F's output gradient = -gy1
    SELFC_LAUNCH_CHECK("cpl_quad0_kernel");
    SELFC_TRY(dense_block_backward<E>(ctx, F, fbuf, ws.fpitch, gyF, 4, gdense, scratch, gF, d, st));
    cpl_rev_pre_kernel<<<nbq, 256, 0, st>>>(gz, gdense, ws.fpitch, gsl, zo, sbuf, gyG, gyH, M);
    SELFC_LAUNCH_CHECK("cpl_rev_pre_kernel");
    SELFC_TRY(dense_block_backward<E>(ctx, G, gbuf, ws.gpitch, gyG, kHF, gdense, scratch, gG, d, st));
    take4_kernel<<<nb, 256, 0, st>>>(acc, gdense, ws.gpitch, gsl, 1, 1.0f, M);
    SELFC_TRY(dense_block_backward<E>(ctx, H, hbuf, ws.gpitch, gyH, kHF, gdense, scratch, gH, d, st));
    take4_kernel<<<nb, 256, 0, st>>>(acc, gdense, ws.gpitch, gsl, 0, 1.0f, M);
    cpl_quad0_kernel<<<nb, 256, 0, st>>>(gz, acc, nullptr, 1.0f, M);
    SELFC_LAUNCH_CHECK("cpl_quad0_kernel");
  }
  return 0;
}

// ====================================================================================================================
// Extra device memory of the training step ("tape"): gradient state, saved block inputs, STP stage outputs, scratch.
// ====================================================================================================================
struct Tape {
  size_t total = 0;
  size_t gz;            // [13][M][4] gradient of the latent state
  size_t zsave;         // 18 x [13][M][4]: state before FA^-1 ... (see train_grads)
  size_t ga;            // 7 x [M][64]: input of STP stage i (i = 0: LR in 4 channels of the first slot) and the final feature
  size_t sa, sb;        // [M][256] scratch each
  size_t sc;            // [M][64] scratch
  size_t gcur, gnext, feat;   // [M][64] each: running feature gradient, its successor, a recomputed STP stage output
  size_t rec;           // [B*T,3,H,W] reconstruction (NCHW), lrq: [B*T,3,h,w] quantised LR (NCHW), glr: [M][4]
  size_t lrq, glr;
  size_t loss;          // 4 floats: sum of squared LR errors, sum of Charbonnier terms
  size_t small;         // per-clip scratch of the GlobalAgg backward
  // the coupling blocks' dense buffers and log-scales, kept by the forward passes so that the backward does not recompute them
  // (16 x (176 + 144 + 144 + 48) x 4 bytes = 32 KB per LR pixel: 1.6 GB per 7x256x448 septuplet -- 180 GB of HBM make the trade easy)
  size_t acts;
  size_t act_f, act_g, act_s, act_block;      // byte sizes of one F buffer, one G / H buffer, one log-scale, one block's set
};
static Tape make_tape(int B, int T, int h, int w) {
  Tape t;
  const size_t M = (size_t)B * T * h * w;
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off = align_up(off + bytes + 256, 1024); return o; };
  t.gz = take(M * kZQuads * 16);
  t.zsave = take(18 * M * kZQuads * 16);
  t.ga = take(7 * M * 64 * 4);
  t.sa = take(M * 256 * 4);
  t.sb = take(M * 256 * 4);
  t.sc = take(M * 64 * 4);
  t.gcur = take(M * 64 * 4);
  t.gnext = take(M * 64 * 4);
  t.feat = take(M * 64 * 4);
  t.rec = take(M * 48 * 4);
  t.lrq = take(M * 3 * 4);
  t.glr = take(M * 4 * 4);
  t.loss = take(64);
  t.small = take(((size_t)B * T * 128 + (size_t)B * T * T + (size_t)h * w + 1024) * 4);
  t.act_f = align_up(M * 176 * 4 + 4096, 1024);
  t.act_g = align_up(M * 144 * 4 + 4096, 1024);
  t.act_s = align_up(M * kHF * 4 + 256, 1024);
  t.act_block = t.act_f + 2 * t.act_g + t.act_s;
  t.acts = take(16 * t.act_block);
  t.total = off;
  return t;
}

// g[m][c] *= (y[m][c] > 0 ? 1 : 0.2) over a whole [M][C] array (C % 4 == 0); y is the activation's OUTPUT (same sign as its input)
__global__ void lrelu_bwd_full_kernel(float* __restrict__ g, const float* __restrict__ y, long long n4) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n4) return;
  float4 gv = load4(g + 4 * i);
  const float4 yv = load4(y + 4 * i);
  gv.x *= yv.x > 0.f ? 1.f : 0.2f; gv.y *= yv.y > 0.f ? 1.f : 0.2f; gv.z *= yv.z > 0.f ? 1.f : 0.2f; gv.w *= yv.w > 0.f ? 1.f : 0.2f;
  store4(g + 4 * i, gv);
}
__global__ void lrelu_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, long long n4) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n4) return;
  const float4 v = load4(x + 4 * i);
  store4(y + 4 * i, make_float4(lrelu02(v.x), lrelu02(v.y), lrelu02(v.z), lrelu02(v.w)));
}

// Backward of a pointwise (1x1x1) conv y = W.x + b:  g_in = W^T g (stored or accumulated),  dW += g (x) in,  db += sum g.
// wf: forward pack [cin][np]; gw / gb: reference layouts [cout][cin] / [cout], accumulated into.
static int pointwise_backward(const selfc_ctx* ctx, const float* wf, int cin, int np, int cout, const float* in, int in_pitch, const float* g,
                              int g_pitch, float* gin, int gin_pitch, bool accumulate, float* gw, float* gb, float* scratch, float* zero_bias,
                              const Dims& d, cudaStream_t st) {
  const long long M = d.M();
  const int cout4 = (cout + 3) & ~3;
  const int npd = (cin + 31) & ~31;
  float* wd = scratch;
  float* dw = scratch + (size_t)cout4 * npd;
  if (gw != nullptr && ctx->mode == SELFC_MODE_BF16X3 && wgrad_tc_on() && cin % 4 == 0 && in_pitch % 4 == 0 && cin <= kWgRows - 1) {
    // BF16X3 mode: the pointwise form of the tensor-core weight-gradient kernel (one unshifted tile per stage), 64 output channels per
    // launch over the same activation planes
    const WgGeom geom = wg_geometry(d);
    void* planes = train_wg_planes(ctx, d, geom, st);
    SELFC_CHECK_ARG(planes != nullptr, "out of device memory (weight-gradient planes, %zu bytes)", wg_plane_bytes(geom));
    SELFC_CUDA(cudaMemsetAsync(dw, 0, (size_t)(cin + 1) * np * sizeof(float), st));
    SELFC_TRY(launch_wg_planes_act_f32(in, in_pitch, cin, d, geom, planes, st));
    for (int n0 = 0; n0 < cout; n0 += 64) {
      const int ncols = cout - n0 < 64 ? cout - n0 : 64;
      const int nb = (ncols + 15) & ~15;
      SELFC_TRY(launch_wg_planes_grad(g, g_pitch, n0, 0, ncols, nb, true, d, geom, planes, st));
      SELFC_TRY(launch_wgrad_tc(planes, geom, cin, ncols, nb, WG_POINT, dw, np, st, n0));
    }
    const long long total = (long long)cout * cin;
    wgrad_unpack_kernel<<<cdiv(total, 256), 256, 0, st>>>(dw, gw, gb, cout, cin, 1, cin, cin, cin, np);
    SELFC_LAUNCH_CHECK("wgrad_unpack_kernel");
  } else if (gw != nullptr) {
    SELFC_CUDA(cudaMemsetAsync(dw, 0, (size_t)(cin + 1) * np * sizeof(float), st));
    const int splits = wgrad_splits(M, 2 * cdiv(cout, 32) * cdiv(cin, WG_C));
    dim3 grid(2 * cdiv(cout, 32), cdiv(cin, WG_C), splits);
    wgrad_kernel<float><<<grid, 256, 0, st>>>(in, in_pitch, 0, cin, g, g_pitch, 0, cout, dw, np, 1, TAP_POINT, d.B * d.T, d.T, d.h, d.w);
    SELFC_LAUNCH_CHECK("wgrad_kernel");
    const long long total = (long long)cout * cin;
    wgrad_unpack_kernel<<<cdiv(total, 256), 256, 0, st>>>(dw, gw, gb, cout, cin, 1, cin, cin, cin, np);
    SELFC_LAUNCH_CHECK("wgrad_unpack_kernel");
  }
  if (gin != nullptr) {
    pack_dgrad_kernel<<<cdiv((long long)cout4 * npd, 256), 256, 0, st>>>(wf, wd, 1, cin, np, cout, cout4, npd);
    SELFC_LAUNCH_CHECK("pack_dgrad_kernel");
    ConvArgs<float> a;
    a.in = g; a.in_pitch = g_pitch; a.cin = cout4;
    a.w = wd; a.bias = zero_bias; a.np = npd; a.cout = cin;
    a.taps = 1; a.tap_mode = TAP_POINT;
    a.BT = d.B * d.T; a.Tn = d.T; a.h = d.h; a.w_ = d.w;
    a.epi = accumulate ? EPI_ACCUM : EPI_STORE; a.outF = gin; a.outF_pitch = gin_pitch; a.outF_off = 0;
    SELFC_TRY(launch_conv_simt<float>(a, st));
  }
  return 0;
}

// ---- soft-GMM sampler backward (SelfC_GMM_arch_inv.py:383-394), one warp per pixel, IN PLACE on the parameters ----
// params [M][720] (channel hf*15 + k*3 + j; j: 0 logit, 1 log-scale, 2 mean) is overwritten with its gradient:
//   g_mu = gv*pi,  g_ls = gv*pi*eps*exp(ls) inside the clamp (0 outside),  g_logit = pi*(a - sum_hf pi*a) with a = gv*(eps*exp(ls)+mu)
__global__ void __launch_bounds__(128) gmm_sample_bwd_kernel(float* __restrict__ params, const float* __restrict__ eps, uint64_t seed,
                                                             uint64_t offset, const float* __restrict__ gz, int T, long long hw, long long M) {
  __shared__ __align__(16) float sp[4][720];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long m = (long long)blockIdx.x * 4 + warp;
  if (m >= M) return;
  const long long n = m / hw, pix = m - n * hw;
  const int t = (int)(n % T);
  const long long b = n / T;
  float* P = sp[warp];
  float4* dst = reinterpret_cast<float4*>(params + m * 720);
  for (int c = lane; c < 180; c += 32) reinterpret_cast<float4*>(P)[c] = dst[c];
  __syncwarp();
  const int hf0 = lane, hf1 = lane + 32;
  const bool has1 = hf1 < kHF;
  const float gv0 = gz[quad_off((size_t)M, 1 + hf0 / 4, (size_t)m) + (hf0 & 3)];
  const float gv1 = has1 ? gz[quad_off((size_t)M, 1 + hf1 / 4, (size_t)m) + (hf1 & 3)] : 0.f;
#pragma unroll
  for (int k = 0; k < kGmmK; ++k) {
    const float l0 = P[hf0 * 15 + k * 3], l1 = has1 ? P[hf1 * 15 + k * 3] : -INFINITY;
    float mx = fmaxf(l0, l1);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    const float e0 = expf(l0 - mx), e1 = has1 ? expf(l1 - mx) : 0.f;
    float sum = e0 + e1;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    const float pi0 = e0 / sum, pi1 = e1 / sum;
    float a0, a1 = 0.f, gls0, gls1 = 0.f;
    {
      const float lsr = P[hf0 * 15 + k * 3 + 1];
      const float sd = expf(fminf(fmaxf(lsr, -7.f), 7.f));
      const float ep = eps ? __ldg(eps + (uint64_t)(((((b * kHF + hf0) * kGmmK + k) * T + t) * hw) + pix))
                           : philox_eps(b, hf0, k, t, pix, T, hw, seed, offset);
      a0 = gv0 * fmaf(ep, sd, P[hf0 * 15 + k * 3 + 2]);
      gls0 = (lsr >= -7.f && lsr <= 7.f) ? gv0 * pi0 * ep * sd : 0.f;
    }
    if (has1) {
      const float lsr = P[hf1 * 15 + k * 3 + 1];
      const float sd = expf(fminf(fmaxf(lsr, -7.f), 7.f));
      const float ep = eps ? __ldg(eps + (uint64_t)(((((b * kHF + hf1) * kGmmK + k) * T + t) * hw) + pix))
                           : philox_eps(b, hf1, k, t, pix, T, hw, seed, offset);
      a1 = gv1 * fmaf(ep, sd, P[hf1 * 15 + k * 3 + 2]);
      gls1 = (lsr >= -7.f && lsr <= 7.f) ? gv1 * pi1 * ep * sd : 0.f;
    }
    float dot = pi0 * a0 + pi1 * a1;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, o);
    __syncwarp();
    P[hf0 * 15 + k * 3] = pi0 * (a0 - dot);
    P[hf0 * 15 + k * 3 + 1] = gls0;
    P[hf0 * 15 + k * 3 + 2] = gv0 * pi0;
    if (has1) {
      P[hf1 * 15 + k * 3] = pi1 * (a1 - dot);
      P[hf1 * 15 + k * 3 + 1] = gls1;
      P[hf1 * 15 + k * 3 + 2] = gv1 * pi1;
    }
  }
  __syncwarp();
  for (int c = lane; c < 180; c += 32) dst[c] = reinterpret_cast<float4*>(P)[c];
}


// GMM head + sampler backward.  feat [M][64]: the STP feature (before the head's first LeakyReLU); gz: gradient state whose HF
// quads hold g_v; gfeat [M][64] receives the feature gradient (overwritten); gparams[6]: tail_gmm.{1,3,5}.{weight,bias}.
// fp32 pixel-major [M][C] -> (hi, lo) pixel-major [M][C] (C % 16 == 0): the input of a tensor-core pointwise conv
__global__ void f32_to_x2_kernel(const float* __restrict__ src, bfx2* __restrict__ dst, long long n4) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n4) store4(dst + 4 * i, load4(src + 4 * i));
}
static int to_x2(const float* src, bfx2* dst, long long M, int C, cudaStream_t st) {
  f32_to_x2_kernel<<<cdiv(M * (C / 4), 256), 256, 0, st>>>(src, dst, M * (C / 4));
  SELFC_LAUNCH_CHECK("f32_to_x2_kernel");
  return 0;
}
// one tensor-core pointwise conv over all M pixels (two pseudo-frames): (hi, lo) input, optional (hi, lo) and / or fp32 outputs
static int pointwise_tc(const TcTempW& w, const bfx2* in, int in_pitch, bfx2* outT, int outT_pitch, float* outF, int outF_pitch, int outF_off,
                        int act, long long M, cudaStream_t st) {
  TcTempArgs t;
  t.in = reinterpret_cast<const __nv_bfloat16*>(in); t.in_pitch = in_pitch; t.B = 1; t.T = 2; t.hw = (int)((M + 1) / 2); t.m_limit = M;
  t.epi = EPI_STORE; t.act = act;
  t.outT = reinterpret_cast<__nv_bfloat16*>(outT); t.outT_pitch = outT_pitch;
  t.outF = outF; t.outF_pitch = outF_pitch; t.outF_off = outF_off;
  return launch_temporal_tc(w, t, st);
}
static bool head_tc_on() { static int c = -1; return env_on("SELFC_HEAD_TC", c); }

int head_sampler_backward(const selfc_ctx* ctx, const float* feat, const float* eps, uint64_t seed, uint64_t offset, const float* gz,
                          float* gfeat, float* const* gparams, char* wsp, const Workspace& ws, char* tp, const Tape& tape, const Dims& d,
                          cudaStream_t st) {
  const long long M = d.M();
  float* h1 = reinterpret_cast<float*>(wsp + ws.h1);
  float* h2 = reinterpret_cast<float*>(wsp + ws.h2);
  float* params = reinterpret_cast<float*>(wsp + ws.params);
  float* fact = reinterpret_cast<float*>(tp + tape.sc);           // [M][64]
  float* gh2 = reinterpret_cast<float*>(tp + tape.sa);            // [M][256]
  float* gh1 = reinterpret_cast<float*>(tp + tape.sb);            // [M][128]
  float* scratch = train_scratch(ctx);
  float* zb = train_zero_bias(ctx);
  SELFC_CHECK_ARG(scratch && zb, "out of device memory (training scratch)");
  const HeadW& hd = ctx->head;
  // BF16X3 mode: the head's recomputed forward and its input gradients as tensor-core pointwise convs ((hi, lo) copies of the fp32
  // tensors in a scratch; fp32 outputs for the masks, the sampler's backward and the weight gradients).  SELFC_HEAD_TC=0: fp32-FMA.
  const bool tc = ctx->mode == SELFC_MODE_BF16X3 && head_tc_on() && hd.r[0].img != nullptr && hd.dg1.img != nullptr;
  bfx2* xs = tc ? train_gslab(ctx, M) : nullptr;
  SELFC_CHECK_ARG(!tc || xs != nullptr, "out of device memory (head scratch)");
  bfx2* fact_x = xs;                                   // [M][64]
  bfx2* h1_x = tc ? fact_x + (size_t)M * 64 : nullptr;   // [M][128]
  bfx2* h2_x = tc ? h1_x + (size_t)M * 128 : nullptr;    // [M][256]
  bfx2* g_x = tc ? h2_x + (size_t)M * 256 : nullptr;     // [M][720]   gradient of the parameters
  bfx2* gh2_x = tc ? g_x + (size_t)M * 720 : nullptr;    // [M][256]
  bfx2* gh1_x = tc ? gh2_x + (size_t)M * 256 : nullptr;  // [M][128]
  // recompute the head: fact = lrelu(feat); h1 = lrelu(W1 fact + b1); h2 = lrelu(W2 h1 + b2); params = W3 h2 + b3
  lrelu_fwd_kernel<<<cdiv(M * 16, 256), 256, 0, st>>>(feat, fact, M * 16);
  SELFC_LAUNCH_CHECK("lrelu_fwd_kernel");
  if (tc) {
    SELFC_TRY(to_x2(fact, fact_x, M, 64, st));
    SELFC_TRY(pointwise_tc(hd.t[0], fact_x, 64, h1_x, 128, h1, 128, 0, 1, M, st));
    SELFC_TRY(pointwise_tc(hd.t[1], h1_x, 128, h2_x, 256, h2, 256, 0, 1, M, st));
    for (int i = 0; i < 5; ++i) SELFC_TRY(pointwise_tc(hd.r[i], h2_x, 256, nullptr, 0, params, 720, 144 * i, 0, M, st));
    gmm_sample_bwd_kernel<<<cdiv(M, 4), 128, 0, st>>>(params, eps, seed, offset, gz, d.T, d.hw(), M);
    SELFC_LAUNCH_CHECK("gmm_sample_bwd_kernel");
    // weight gradients through pointwise_backward (tensor-core form, no input gradient: gin = null), input gradients here
    SELFC_TRY(pointwise_backward(ctx, hd.w[2], 256, hd.np[2], 720, h2, 256, params, 720, nullptr, 256, false, gparams ? gparams[4] : nullptr,
                                 gparams ? gparams[5] : nullptr, scratch, zb, d, st));
    SELFC_TRY(to_x2(params, g_x, M, 720, st));
    for (int j = 0; j < 4; ++j) SELFC_TRY(pointwise_tc(hd.dg3[j], g_x, 720, nullptr, 0, gh2, 256, 64 * j, 0, M, st));
    lrelu_bwd_full_kernel<<<cdiv(M * 64, 256), 256, 0, st>>>(gh2, h2, M * 64);
    SELFC_TRY(pointwise_backward(ctx, hd.w[1], 128, hd.np[1], 256, h1, 128, gh2, 256, nullptr, 128, false, gparams ? gparams[2] : nullptr,
                                 gparams ? gparams[3] : nullptr, scratch, zb, d, st));
    SELFC_TRY(to_x2(gh2, gh2_x, M, 256, st));
    SELFC_TRY(pointwise_tc(hd.dg2, gh2_x, 256, nullptr, 0, gh1, 128, 0, 0, M, st));
    lrelu_bwd_full_kernel<<<cdiv(M * 32, 256), 256, 0, st>>>(gh1, h1, M * 32);
    SELFC_TRY(pointwise_backward(ctx, hd.w[0], 64, hd.np[0], 128, fact, 64, gh1, 128, nullptr, 64, false, gparams ? gparams[0] : nullptr,
                                 gparams ? gparams[1] : nullptr, scratch, zb, d, st));
    SELFC_TRY(to_x2(gh1, gh1_x, M, 128, st));
    SELFC_TRY(pointwise_tc(hd.dg1, gh1_x, 128, nullptr, 0, gfeat, 64, 0, 0, M, st));
    lrelu_bwd_full_kernel<<<cdiv(M * 16, 256), 256, 0, st>>>(gfeat, fact, M * 16);
    SELFC_LAUNCH_CHECK("lrelu_bwd_full_kernel");
    return 0;
  }
  ConvArgs<float> a;
  a.BT = d.B * d.T; a.Tn = d.T; a.h = d.h; a.w_ = d.w;
  a.taps = 1; a.tap_mode = TAP_POINT; a.epi = EPI_STORE;
  a.in = fact; a.in_pitch = 64; a.cin = 64; a.w = hd.w[0]; a.bias = hd.b[0]; a.np = hd.np[0]; a.cout = 128; a.act = 1; a.outF = h1; a.outF_pitch = 128;
  SELFC_TRY(launch_conv_simt<float>(a, st));
  a.in = h1; a.in_pitch = 128; a.cin = 128; a.w = hd.w[1]; a.bias = hd.b[1]; a.np = hd.np[1]; a.cout = 256; a.outF = h2; a.outF_pitch = 256;
  SELFC_TRY(launch_conv_simt<float>(a, st));
  a.in = h2; a.in_pitch = 256; a.cin = 256; a.w = hd.w[2]; a.bias = hd.b[2]; a.np = hd.np[2]; a.cout = 720; a.act = 0; a.outF = params; a.outF_pitch = 720;
  SELFC_TRY(launch_conv_simt<float>(a, st));
  // sampler backward, in place: params <- d loss / d params
  gmm_sample_bwd_kernel<<<cdiv(M, 4), 128, 0, st>>>(params, eps, seed, offset, gz, d.T, d.hw(), M);
  SELFC_LAUNCH_CHECK("gmm_sample_bwd_kernel");
  // 256 -> 720
  SELFC_TRY(pointwise_backward(ctx, hd.w[2], 256, hd.np[2], 720, h2, 256, params, 720, gh2, 256, false, gparams ? gparams[4] : nullptr,
                               gparams ? gparams[5] : nullptr, scratch, zb, d, st));
  lrelu_bwd_full_kernel<<<cdiv(M * 64, 256), 256, 0, st>>>(gh2, h2, M * 64);
  // 128 -> 256
  SELFC_TRY(pointwise_backward(ctx, hd.w[1], 128, hd.np[1], 256, h1, 128, gh2, 256, gh1, 128, false, gparams ? gparams[2] : nullptr,
                               gparams ? gparams[3] : nullptr, scratch, zb, d, st));
  lrelu_bwd_full_kernel<<<cdiv(M * 32, 256), 256, 0, st>>>(gh1, h1, M * 32);
  // 64 -> 128
  SELFC_TRY(pointwise_backward(ctx, hd.w[0], 64, hd.np[0], 128, fact, 64, gh1, 128, gfeat, 64, false, gparams ? gparams[0] : nullptr,
                               gparams ? gparams[1] : nullptr, scratch, zb, d, st));
  lrelu_bwd_full_kernel<<<cdiv(M * 16, 256), 256, 0, st>>>(gfeat, fact, M * 16);
  SELFC_LAUNCH_CHECK("lrelu_bwd_full_kernel");
  return 0;
}

// ====================================================================================================================
// GlobalAgg backward (SelfC_GMM_arch_inv.py:265-285).  Forward (as computed here): d = fc(pool(x)) = fcb + sum_pix wmap*x,
// q = P2 d + b2, k = P3 d + b3, W = softmax_rows(q k^T / 64), Xmix[t'] = sum_t W[t,t'] x[t],
// out[t'] = x[t'] + P1 Xmix[t'] + b1 * sum_t W[t,t'].
// ====================================================================================================================
// out[b,t'] = (base ? base : 0) + sum_t W[b,t,t'] in[b,t]   (transpose: out[b,t] = sum_t' W[b,t,t'] in[b,t'])
__global__ void ga_mix_f32_kernel(const float* __restrict__ in, const float* __restrict__ wmat, int transpose, const float* __restrict__ base,
                                  float* __restrict__ out, int T, long long hw, long long M) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long m = idx >> 4;
  const int c = (int)(idx & 15) * 4;
  if (m >= M) return;
  const long long n = m / hw, pix = m - n * hw;
  const int to = (int)(n % T);
  const long long b = n / T;
  float4 acc = base ? load4(base + m * 64 + c) : make_float4(0.f, 0.f, 0.f, 0.f);
  const float* wm = wmat + b * T * T;
  for (int t = 0; t < T; ++t) {
    const float wv = transpose ? __ldg(wm + to * T + t) : __ldg(wm + t * T + to);
    const float4 v = load4(in + ((b * T + t) * hw + pix) * 64 + c);
    acc.x = fmaf(wv, v.x, acc.x); acc.y = fmaf(wv, v.y, acc.y); acc.z = fmaf(wv, v.z, acc.z); acc.w = fmaf(wv, v.w, acc.w);
  }
  store4(out + m * 64 + c, acc);
}

// Both reductions below run as (output, pixel split) blocks writing partial sums part[output * nsplit + split], summed in a fixed
// order by sum_splits_kernel (deterministic): one block per output left 7 / 49 CTAs on 148 SMs (58 / 128 us at Vimeo shape).
constexpr int kGaSplits = 16;
__global__ void sum_splits_kernel(const float* __restrict__ part, float* __restrict__ out, int n_out, int nsplit) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_out) return;
  float s = 0.f;
  for (int k = 0; k < nsplit; ++k) s += part[(size_t)i * nsplit + k];
  out[i] = s;
}
// S[n][ch] = sum_pix g[n,pix,ch]   (grid (frames, splits), 256 threads = 4 pixel lanes x 64 channels)
__global__ void __launch_bounds__(256) frame_channel_sum_kernel(const float* __restrict__ g, float* __restrict__ Spart, long long hw) {
  __shared__ float red[4][64];
  const int n = blockIdx.x, split = blockIdx.y, nsplit = gridDim.y, ch = threadIdx.x & 63, lane = threadIdx.x >> 6;
  const long long per = (hw + nsplit - 1) / nsplit, p0 = split * per, p1 = p0 + per < hw ? p0 + per : hw;
  float acc = 0.f;
  for (long long p = p0 + lane; p < p1; p += 4) acc += g[((long long)n * hw + p) * 64 + ch];
  red[lane][ch] = acc;
  __syncthreads();
  if (lane == 0) Spart[((size_t)n * 64 + ch) * nsplit + split] = red[0][ch] + red[1][ch] + red[2][ch] + red[3][ch];
}

// R[b][t][u] = sum_{pix,ch} a[b,t,pix,ch] * c[b,u,pix,ch]   (grid ((b,t,u), splits))
__global__ void __launch_bounds__(256) pair_dot_kernel(const float* __restrict__ a, const float* __restrict__ c, float* __restrict__ Rpart, int T,
                                                       long long hw) {
  __shared__ float red[256];
  const int u = blockIdx.x % T, t = (blockIdx.x / T) % T, b = blockIdx.x / (T * T);
  const int split = blockIdx.y, nsplit = gridDim.y;
  const float4* pa = reinterpret_cast<const float4*>(a + ((long long)(b * T + t) * hw) * 64);
  const float4* pc = reinterpret_cast<const float4*>(c + ((long long)(b * T + u) * hw) * 64);
  const long long n4 = hw * 16, per = (n4 + nsplit - 1) / nsplit, i0 = split * per, i1 = i0 + per < n4 ? i0 + per : n4;
  float acc = 0.f;
  for (long long i = i0 + threadIdx.x; i < i1; i += 256) {
    const float4 x = pa[i], y = pc[i];
    acc += x.x * y.x + x.y * y.y + x.z * y.z + x.w * y.w;
  }
  red[threadIdx.x] = acc;
  __syncthreads();
  for (int s2 = 128; s2 > 0; s2 >>= 1) {
    if (threadIdx.x < s2) red[threadIdx.x] += red[threadIdx.x + s2];
    __syncthreads();
  }
  if (threadIdx.x == 0) Rpart[(size_t)blockIdx.x * nsplit + split] = red[0];
}

// per clip: recompute d, q, k, W; gW = R + b1 . S; softmax backward; gq, gk; gd; gradients of proj2 / proj3 / fc.bias
__global__ void __launch_bounds__(64) ga_small_bwd_kernel(const float* __restrict__ partial, int nsplit, const float* __restrict__ fcb,
                                                          const float* __restrict__ p2w, const float* __restrict__ p2b,
                                                          const float* __restrict__ p3w, const float* __restrict__ p3b,
                                                          const float* __restrict__ b1, const float* __restrict__ R,
                                                          const float* __restrict__ S, float* __restrict__ gd, float* __restrict__ gp2w,
                                                          float* __restrict__ gp2b, float* __restrict__ gp3w, float* __restrict__ gp3b,
                                                          float* __restrict__ gfcb, int T) {
  constexpr int MAXT = 32;
  extern __shared__ float sm[];
  float* d = sm;                 // [T][64]
  float* q = d + T * 64;         // [T][64]
  float* k = q + T * 64;         // [T][64]
  float* gq = k + T * 64;        // [T][64]
  float* gk = gq + T * 64;       // [T][64]
  float* W = gk + T * 64;        // [T][MAXT]
  float* gA = W + T * MAXT;      // [T][MAXT]
  const int b = blockIdx.x, c = threadIdx.x;
  for (int t = 0; t < T; ++t) {
    float s = 0.f;
    for (int sp = 0; sp < nsplit; ++sp) s += partial[(((long long)b * T + t) * nsplit + sp) * 64 + c];
    d[t * 64 + c] = s + fcb[0];
  }
  __syncthreads();
  for (int t = 0; t < T; ++t) {
    float sq = p2b[c], sk = p3b[c];
    for (int i = 0; i < 64; ++i) {
      sq += p2w[c * 64 + i] * d[t * 64 + i];
      sk += p3w[c * 64 + i] * d[t * 64 + i];
    }
    q[t * 64 + c] = sq;
    k[t * 64 + c] = sk;
  }
  __syncthreads();
  for (int e = c; e < T * T; e += 64) {
    const int t = e / T, u = e % T;
    float s = 0.f;
    for (int i = 0; i < 64; ++i) s += q[t * 64 + i] * k[u * 64 + i];
    W[t * MAXT + u] = s / 64.0f;
    float gw = R[((long long)b * T + t) * T + u];
    for (int i = 0; i < 64; ++i) gw += b1[i] * S[((long long)b * T + u) * 64 + i];
    gA[t * MAXT + u] = gw;                    // holds gW until the softmax backward below
  }
  __syncthreads();
  if (c < T) {
    const int t = c;
    float mx = -INFINITY;
    for (int u = 0; u < T; ++u) mx = fmaxf(mx, W[t * MAXT + u]);
    float sum = 0.f;
    for (int u = 0; u < T; ++u) { W[t * MAXT + u] = expf(W[t * MAXT + u] - mx); sum += W[t * MAXT + u]; }
    float dot = 0.f;
    for (int u = 0; u < T; ++u) { W[t * MAXT + u] /= sum; dot += W[t * MAXT + u] * gA[t * MAXT + u]; }
    for (int u = 0; u < T; ++u) gA[t * MAXT + u] = W[t * MAXT + u] * (gA[t * MAXT + u] - dot) / 64.0f;   // d loss / d (q.k), 1/64 folded in
  }
  __syncthreads();
  for (int t = 0; t < T; ++t) {
    float a = 0.f, bb = 0.f;
    for (int u = 0; u < T; ++u) {
      a += gA[t * MAXT + u] * k[u * 64 + c];       // gq[t][c]
      bb += gA[u * MAXT + t] * q[u * 64 + c];      // gk[t][c]
    }
    gq[t * 64 + c] = a;
    gk[t * 64 + c] = bb;
  }
  __syncthreads();
  float sfc = 0.f;
  for (int t = 0; t < T; ++t) {
    float g = 0.f;                                  // gd[t][i = c]
    for (int o = 0; o < 64; ++o) g += p2w[o * 64 + c] * gq[t * 64 + o] + p3w[o * 64 + c] * gk[t * 64 + o];
    gd[((long long)b * T + t) * 64 + c] = g;
    sfc += g;
  }
  // parameter gradients (row c of proj2 / proj3)
  if (gp2w) {
    float sb2 = 0.f, sb3 = 0.f;
    for (int t = 0; t < T; ++t) { sb2 += gq[t * 64 + c]; sb3 += gk[t * 64 + c]; }
    atomicAdd(gp2b + c, sb2);
    atomicAdd(gp3b + c, sb3);
    for (int i = 0; i < 64; ++i) {
      float a2 = 0.f, a3 = 0.f;
      for (int t = 0; t < T; ++t) { a2 += gq[t * 64 + c] * d[t * 64 + i]; a3 += gk[t * 64 + c] * d[t * 64 + i]; }
      atomicAdd(gp2w + c * 64 + i, a2);
      atomicAdd(gp3w + c * 64 + i, a3);
    }
    // fc.bias: sum of gd over frames and channels
    __shared__ float red[64];
    red[c] = sfc;
    __syncthreads();
    if (c == 0) {
      float s = 0.f;
      for (int i = 0; i < 64; ++i) s += red[i];
      atomicAdd(gfcb, s);
    }
  }
}

// descriptor path: gx[n,pix,:] += wmap[pix] * gd[n,:];  gwmap[pix] += sum_ch gd[n,ch] * x[n,pix,ch]
__global__ void ga_desc_bwd_kernel(const float* __restrict__ x, const float* __restrict__ wmap, const float* __restrict__ gd,
                                   float* __restrict__ gx, float* __restrict__ gwmap, long long hw, long long M) {
  const long long m = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= M) return;
  const long long n = m / hw, pix = m - n * hw;
  const float wv = wmap[pix];
  float dot = 0.f;
  for (int c = 0; c < 64; c += 4) {
    const float4 g = load4(gd + n * 64 + c);
    const float4 xv = load4(x + m * 64 + c);
    float4 o = load4(gx + m * 64 + c);
    o.x += wv * g.x; o.y += wv * g.y; o.z += wv * g.z; o.w += wv * g.w;
    store4(gx + m * 64 + c, o);
    dot += g.x * xv.x + g.y * xv.y + g.z * xv.z + g.w * xv.w;
  }
  atomicAdd(gwmap + pix, dot);
}

// adjoint of ga_wmap_kernel: gfcw[i*32+j] += sum over the pixels of pooling bin (i,j) of gwmap / |bin|
__global__ void ga_wmap_bwd_kernel(const float* __restrict__ gwmap, float* __restrict__ gfcw, int h, int w) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= 1024) return;
  const int i = e / 32, j = e % 32;
  const int ys = (i * h) / 32, ye = ((i + 1) * h + 31) / 32;
  const int xs = (j * w) / 32, xe = ((j + 1) * w + 31) / 32;
  float acc = 0.f;
  for (int y = ys; y < ye; ++y)
    for (int x = xs; x < xe; ++x) acc += gwmap[y * w + x];
  gfcw[e] += acc / (float)((ye - ys) * (xe - xs));
}

__global__ void ga_bias1_kernel(const float* __restrict__ S, const float* __restrict__ wsum, float* __restrict__ gb1, int BT) {
  const int c = threadIdx.x;
  float acc = 0.f;
  for (int n = 0; n < BT; ++n) acc += wsum[n] * S[n * 64 + c];
  gb1[c] += acc;
}

// x, gout: [M][64]; gx [M][64] out; gparams[8]: fc.weight, fc.bias, proj1.weight, proj1.bias, proj2.weight, proj2.bias, proj3.weight, proj3.bias
int ga_backward(const selfc_ctx* ctx, const GaW& g, const float* x, const float* gout, float* gx, float* const* gparams, char* wsp,
                const Workspace& ws, char* tp, const Tape& tape, const Dims& d, cudaStream_t st) {
  const long long M = d.M(), hw = d.hw();
  const int T = d.T, BT = d.B * d.T;
  float* wmap = reinterpret_cast<float*>(wsp + ws.wmap);
  float* partial = reinterpret_cast<float*>(wsp + ws.partial);
  float* wmat = reinterpret_cast<float*>(wsp + ws.wmat);
  float* wsum = reinterpret_cast<float*>(wsp + ws.wsum);
  float* xmix = reinterpret_cast<float*>(tp + tape.sc);          // [M][64]
  float* gxmix = reinterpret_cast<float*>(tp + tape.sa);         // [M][64]
  float* small = reinterpret_cast<float*>(tp + tape.small);
  auto up64 = [](size_t n) { return (n + 63) & ~(size_t)63; };   // keep every sub-buffer 16-byte aligned
  float* S = small;                                  // [BT][64]
  float* R = S + (size_t)BT * 64;                    // [B][T][T]
  float* gd = R + up64((size_t)d.B * T * T);         // [BT][64]
  float* gwmap = gd + (size_t)BT * 64;               // [hw]
  float* gb1w = gwmap + up64((size_t)hw);            // [64] scratch for the wsum-weighted bias gradient
  float* scratch = train_scratch(ctx);
  float* zb = train_zero_bias(ctx);
  SELFC_CHECK_ARG(scratch && zb, "out of device memory (training scratch)");
  // recompute the forward's small quantities
  SELFC_TRY(launch_ga_wmap(g.fcw, wmap, d.h, d.w, st));
  SELFC_TRY(launch_ga_stat<float>(x, kStpC, wmap, partial, ws.nsplit, BT, (int)hw, st));
  SELFC_TRY(launch_ga_weights(partial, ws.nsplit, g.fcb, g.p2w, g.p2b, g.p3w, g.p3b, wmat, wsum, d.B, T, st));
  const int nb16 = cdiv(M * 16, 256);
  ga_mix_f32_kernel<<<nb16, 256, 0, st>>>(x, wmat, 0, nullptr, xmix, T, hw, M);
  SELFC_LAUNCH_CHECK("ga_mix_f32_kernel");
  // proj1: weight gradient from (Xmix, gout); gXmix = P1^T gout.  Its plain bias gradient is NOT the right one (the bias is
  // scaled by colsum(W) per frame), so it goes to a scratch and the weighted sum is formed from the per-frame sums S below.
  SELFC_CUDA(cudaMemsetAsync(gb1w, 0, 64 * sizeof(float), st));
  SELFC_TRY(pointwise_backward(ctx, g.p1w, 64, 64, 64, xmix, 64, gout, 64, gxmix, 64, false, gparams ? gparams[2] : nullptr, gb1w, scratch, zb, d, st));
  // (the training scratch is free again once pointwise_backward's launches are queued: partial sums of the two reductions)
  float* Spart = scratch;
  float* Rpart = scratch + (size_t)BT * 64 * kGaSplits;
  frame_channel_sum_kernel<<<dim3(BT, kGaSplits), 256, 0, st>>>(gout, Spart, hw);
  SELFC_LAUNCH_CHECK("frame_channel_sum_kernel");
  sum_splits_kernel<<<cdiv(BT * 64, 256), 256, 0, st>>>(Spart, S, BT * 64, kGaSplits);
  SELFC_LAUNCH_CHECK("sum_splits_kernel");
  // gx = gout + sum_t' W[t,t'] gXmix[t']
  ga_mix_f32_kernel<<<nb16, 256, 0, st>>>(gxmix, wmat, 1, gout, gx, T, hw, M);
  SELFC_LAUNCH_CHECK("ga_mix_f32_kernel");
  pair_dot_kernel<<<dim3(d.B * T * T, kGaSplits), 256, 0, st>>>(x, gxmix, Rpart, T, hw);
  SELFC_LAUNCH_CHECK("pair_dot_kernel");
  sum_splits_kernel<<<cdiv(d.B * T * T, 256), 256, 0, st>>>(Rpart, R, d.B * T * T, kGaSplits);
  SELFC_LAUNCH_CHECK("sum_splits_kernel");
  SELFC_CUDA(cudaMemsetAsync(gwmap, 0, (size_t)hw * sizeof(float), st));
  const size_t smem = (size_t)(5 * T * 64 + 2 * T * 32) * sizeof(float);
  ga_small_bwd_kernel<<<d.B, 64, smem, st>>>(partial, ws.nsplit, g.fcb, g.p2w, g.p2b, g.p3w, g.p3b, g.p1b, R, S, gd,
                                             gparams ? gparams[4] : nullptr, gparams ? gparams[5] : nullptr, gparams ? gparams[6] : nullptr,
                                             gparams ? gparams[7] : nullptr, gparams ? gparams[1] : nullptr, T);
  SELFC_LAUNCH_CHECK("ga_small_bwd_kernel");
  ga_desc_bwd_kernel<<<cdiv(M, 256), 256, 0, st>>>(x, wmap, gd, gx, gwmap, hw, M);
  SELFC_LAUNCH_CHECK("ga_desc_bwd_kernel");
  if (gparams) {
    ga_wmap_bwd_kernel<<<4, 256, 0, st>>>(gwmap, gparams[0], d.h, d.w);
    SELFC_LAUNCH_CHECK("ga_wmap_bwd_kernel");
    // proj1.bias: sum_{b,t'} wsum[b,t'] * S[b,t',ch]
    ga_bias1_kernel<<<1, 64, 0, st>>>(S, wsum, gparams[3], BT);
    SELFC_LAUNCH_CHECK("ga_bias1_kernel");
  }
  return 0;
}

// ====================================================================================================================
// The training step's forward + backward (models/SelfC_model.py:148-170): losses and all 354 parameter gradients.
// ====================================================================================================================
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// l_forw_fit = mean((LR_pre - ref_L)^2) (loss.py:12-13): accumulates the sum into loss[0] and writes the LR part of the state
// gradient: gz quad 0 = (gq0 from the reverse pass, straight through the quantiser) + glr (from the STP's first block) +
// 2*(lr_pre - ref)*gscale; the HF part of the downscaling output does not enter the loss: gz quads 1..12 = 0.
__global__ void loss_forw_kernel(const float* __restrict__ zout, const float* __restrict__ ref_l, const float* __restrict__ glr,
                                 float* __restrict__ gz, float* __restrict__ loss, float gscale, long long hw, long long M) {
  const long long m = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  float part = 0.f;
  if (m < M) {
    const long long n = m / hw, pix = m - n * hw;
    const float4 lr = load4(zout + quad_off((size_t)M, 0, (size_t)m));
    float4 g = load4(gz + quad_off((size_t)M, 0, (size_t)m));
    const float4 gs = load4(glr + m * 4);
    const float d0 = lr.x - ref_l[(n * 3 + 0) * hw + pix], d1 = lr.y - ref_l[(n * 3 + 1) * hw + pix], d2 = lr.z - ref_l[(n * 3 + 2) * hw + pix];
    part = d0 * d0 + d1 * d1 + d2 * d2;
    g.x += gs.x + 2.f * d0 * gscale; g.y += gs.y + 2.f * d1 * gscale; g.z += gs.z + 2.f * d2 * gscale; g.w = 0.f;
    store4(gz + quad_off((size_t)M, 0, (size_t)m), g);
    for (int q = 1; q < kZQuads; ++q) store4(gz + quad_off((size_t)M, q, (size_t)m), make_float4(0.f, 0.f, 0.f, 0.f));
  }
  part = warp_sum(part);
  if ((threadIdx.x & 31) == 0) atomicAdd(loss, part);
}

// l_back_rec = mean(sqrt((x - rec)^2 + 1e-6)) (loss.py:15-17) and the backward of the reverse FrequencyAnalyzer
// (rec[c,4i+sy,4j+sx] = lf[c] + hf[c*16+sy*4+sx], SelfC_GMM_arch_inv.py:79-82): one thread per LR pixel writes the whole
// gradient state: quad 0 = sum over the 4x4 block, HF quads = the per-pixel gradients in the reverse (c*16+sy*4+sx) order.
__global__ void loss_back_fa_bwd_kernel(const float* __restrict__ x, const float* __restrict__ rec, float* __restrict__ gz,
                                        float* __restrict__ loss, float gscale, int N, int h, int w) {
  const long long M = (long long)N * h * w;
  const long long m = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  float part = 0.f;
  if (m < M) {
    const int j = (int)(m % w);
    const int i = (int)((m / w) % h);
    const int n = (int)(m / ((long long)w * h));
    const int W = 4 * w, H = 4 * h;
    float glf[3] = {0.f, 0.f, 0.f};
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const long long base = (((long long)n * 3 + c) * H + 4 * i) * W + 4 * j;
#pragma unroll
      for (int sy = 0; sy < 4; ++sy) {
        const float4 xv = load4(x + base + (long long)sy * W);
        const float4 rv = load4(rec + base + (long long)sy * W);
        const float dd[4] = {xv.x - rv.x, xv.y - rv.y, xv.z - rv.z, xv.w - rv.w};
        float g4[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float r = sqrtf(dd[e] * dd[e] + 1e-6f);
          part += r;
          g4[e] = -dd[e] / r * gscale;
          glf[c] += g4[e];
        }
        store4(gz + quad_off((size_t)M, 1 + (c * 16 + sy * 4) / 4, (size_t)m), make_float4(g4[0], g4[1], g4[2], g4[3]));
      }
    }
    store4(gz + quad_off((size_t)M, 0, (size_t)m), make_float4(glf[0], glf[1], glf[2], 0.f));
  }
  part = warp_sum(part);
  if ((threadIdx.x & 31) == 0) atomicAdd(loss + 1, part);
}

// dst[m][doff + c] = src[m][soff + c], c < ncol (ncol % 4 == 0)
__global__ void copy_cols_kernel(float* __restrict__ dst, int dpitch, int doff, const float* __restrict__ src, int spitch, int soff, int ncol,
                                 long long M, long long sslabM = 0) {
  const int per = ncol / 4;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= M * per) return;
  const long long m = idx / per;
  const int c = (int)(idx - m * per) * 4;
  store4(dst + m * dpitch + doff + c, load4(src + dense_off(m, soff + c, spitch, sslabM)));
}

// X slot of a dense buffer (either layout) <- fp32 pixel-major [M][spitch] columns [0, ncol)
template <typename E>
__global__ void cols_to_dense_kernel(E* __restrict__ dst, int dpitch, long long slabM, const float* __restrict__ src, int spitch, int ncol,
                                     long long M) {
  const int per = ncol / 4;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= M * per) return;
  const long long m = idx / per;
  const int c = (int)(idx - m * per) * 4;
  store4(dst + dense_off(m, c, dpitch, slabM), load4(src + m * spitch + c));
}

// losses[0..2] = (total, l_forw_fit, l_back_rec) from the accumulated sums
__global__ void loss_finish_kernel(const float* __restrict__ acc, float* __restrict__ out, float n_lr, float n_hr) {
  const float lf = acc[0] / n_lr, lb = acc[1] / n_hr;
  out[0] = (lf + lb) * (144.0f * 144.0f * 3.0f);
  out[1] = lf;
  out[2] = lb;
}

// per-block buffers inside the tape: dir 0 = the downscaling pass (blocks run 0..7), dir 1 = the upscaling pass (7..0); the last
// block of a pass has no successor: its "next" targets are the workspace's own buffers (never read)
static void make_block_bufs(BlockBufsV out[8], int dir, char* tp, const Tape& tape, char* wsp, const Workspace& ws) {
  auto blockp = [&](int blk) { return tp + tape.acts + (size_t)(dir * 8 + blk) * tape.act_block; };
  for (int blk = 0; blk < 8; ++blk) {
    char* b = blockp(blk);
    out[blk].f = b;
    out[blk].g = b + tape.act_f;
    out[blk].h = b + tape.act_f + tape.act_g;
    out[blk].s = reinterpret_cast<float*>(b + tape.act_f + 2 * tape.act_g);
    const int nxt = dir == 0 ? blk + 1 : blk - 1;
    const bool has = nxt >= 0 && nxt < 8;
    out[blk].f_next = has ? (void*)blockp(nxt) : (void*)(wsp + ws.fbuf);
    out[blk].g_next = has ? (void*)(blockp(nxt) + tape.act_f) : (void*)(wsp + ws.gbuf);
    out[blk].h_next = has ? (void*)(blockp(nxt) + tape.act_f + tape.act_g) : (void*)(wsp + ws.hbuf);
  }
}

template <typename E>
int train_grads(selfc_ctx* ctx, const float* hr, const float* ref_l, const float* eps, uint64_t seed, uint64_t offset, float* const* grads,
                float* losses, const Dims& d, char* wsp, const Workspace& ws, char* tp, const Tape& tape, cudaStream_t st) {
  const long long M = d.M(), hw = d.hw();
  const long long slabM = dense_slab(ctx, d);
  const int BT = d.B * d.T;
  const size_t state_f = (size_t)M * kZQuads * 4;       // floats per planar state
  float* z = reinterpret_cast<float*>(wsp + ws.z);
  E* fbuf = reinterpret_cast<E*>(wsp + ws.fbuf);
  E* stpbuf = reinterpret_cast<E*>(wsp + ws.stpbuf);
  float* gdense = reinterpret_cast<float*>(wsp + ws.params);
  float* gz = reinterpret_cast<float*>(tp + tape.gz);
  float* zs_dn = reinterpret_cast<float*>(tp + tape.zsave);             // slots 0..7: input of forward block blk; slot 8: the output
  float* zs_up = zs_dn + 9 * state_f;                                   // slots 0..7: input of reverse block blk
  float* ga_save = reinterpret_cast<float*>(tp + tape.ga);
  float* gcur = reinterpret_cast<float*>(tp + tape.gcur);
  float* gnext = reinterpret_cast<float*>(tp + tape.gnext);
  float* feat = reinterpret_cast<float*>(tp + tape.feat);
  float* rec = reinterpret_cast<float*>(tp + tape.rec);
  float* lrq = reinterpret_cast<float*>(tp + tape.lrq);
  float* glr = reinterpret_cast<float*>(tp + tape.glr);
  float* lacc = reinterpret_cast<float*>(tp + tape.loss);
  float* scratch = train_scratch(ctx);
  SELFC_CHECK_ARG(scratch != nullptr, "out of device memory (training scratch)");
  const float kScale = 144.0f * 144.0f * 3.0f;
  const float n_lr = (float)((double)M * 3.0), n_hr = (float)((double)M * 48.0);
  SELFC_CUDA(cudaMemsetAsync(lacc, 0, 64, st));

  // ---- forward: downscale (states kept), quantise, upscale (states and STP stage outputs kept) ----
  // every coupling block of both passes runs in its own buffers inside the tape and leaves them for its backward (SELFC_TRAIN_KEEP=0:
  // the three workspace buffers, every block's forward recomputed from its saved input state)
  static int keep_on = -1;
  const bool keep = env_on("SELFC_TRAIN_KEEP", keep_on);
  BlockBufsV bufs_dn[8], bufs_up[8];
  make_block_bufs(bufs_dn, 0, tp, tape, wsp, ws);
  make_block_bufs(bufs_up, 1, tp, tape, wsp, ws);
  SELFC_TRY(launch_fa_fwd_z<E>(hr, z, keep ? static_cast<E*>(bufs_dn[0].f) : fbuf, ws.fpitch, slabM, BT, d.h, d.w, st));
  for (int blk = 0; blk < 8; ++blk) {
    SELFC_CUDA(cudaMemcpyAsync(zs_dn + blk * state_f, z, state_f * 4, cudaMemcpyDeviceToDevice, st));
    SELFC_TRY(invblock_fwd<E>(ctx, blk, false, wsp, ws, d, st, keep ? &bufs_dn[blk] : nullptr));
  }
  SELFC_CUDA(cudaMemcpyAsync(zs_dn + 8 * state_f, z, state_f * 4, cudaMemcpyDeviceToDevice, st));
  SELFC_TRY(launch_export_down(z, nullptr, nullptr, lrq, M, hw, st));            // Quantization.forward
  TrainHooks hooks;
  hooks.ga_save = ga_save;
  hooks.z_save = zs_up;
  hooks.up_bufs = keep ? bufs_up : nullptr;
  SELFC_TRY(up_hooked<E>(ctx, lrq, eps, seed, offset, rec, d, wsp, ws, st, &hooks));

  // ---- backward ----
  // input-gradient images of all 30 dense blocks, stale since the weight load of this step: three launches instead of seven per block
  if constexpr (std::is_same<E, bfx2>::value) {
    if (dgrad_tc_on() && pack_batch_enabled()) {
      float* zb = train_zero_bias(ctx);
      SELFC_CHECK_ARG(zb != nullptr, "out of device memory (training scratch)");
      PackBatchScope scope(ctx->pack_dg);
      for (int blk = 0; blk < 8; ++blk)
        for (int j = 0; j < 3; ++j) SELFC_TRY(ensure_dgrad_images(ctx, ctx->inv[blk][j], zb, st));
      for (int i = 0; i < 6; ++i) SELFC_TRY(ensure_dgrad_images(ctx, ctx->stp[i], zb, st));
      SELFC_TRY(pack_batch_flush(ctx->pack_dg, st));
    }
  }
  loss_back_fa_bwd_kernel<<<cdiv(M, 128), 128, 0, st>>>(hr, rec, gz, lacc, kScale / n_hr, BT, d.h, d.w);
  SELFC_LAUNCH_CHECK("loss_back_fa_bwd_kernel");
  for (int blk = 0; blk < 8; ++blk)            // the reverse pass ran blocks 7..0, so their backward runs 0..7
    SELFC_TRY(invblock_backward<E>(ctx, blk, true, zs_up + blk * state_f, gz, grads ? grads + P_INV0 + 30 * blk : nullptr, wsp, ws, d, st,
                                   keep ? &bufs_up[blk] : nullptr, zs_up + (blk == 0 ? 8 : blk - 1) * state_f));
  // gz now holds d loss / d [LR_q | v]; the HF part goes through the sampler and the GMM head into the STP
  SELFC_TRY(head_sampler_backward(ctx, ga_save + (size_t)6 * M * kStpC, eps, seed, offset, gz, gcur, grads ? grads + P_TAIL : nullptr, wsp, ws,
                                  tp, tape, d, st));
  const int stp_first[6] = {P_LOCAL1, P_LOCAL2, P_OTHER, P_OTHER + 18, P_OTHER + 36, P_OTHER + 54};
  const int ga_first[6] = {P_GLOBAL1, P_GLOBAL2, P_OTHER + 10, P_OTHER + 28, P_OTHER + 46, P_OTHER + 64};
  for (int i = 5; i >= 0; --i) {
    const DenseW& W = ctx->stp[i];
    const int pitch = W.xpad + 4 * kGrowth;
    // recompute stage i: X slot <- its input (LR for the first stage), dense block -> feat
    if (i == 0) SELFC_TRY(launch_nchw_to_dense<E>(lrq, stpbuf, pitch, slabM, 0, 3, W.xpad, M, hw, st));
    else {
      cols_to_dense_kernel<E><<<cdiv(M * 16, 256), 256, 0, st>>>(stpbuf, pitch, slabM, ga_save + (size_t)i * M * kStpC, kStpC, kStpC, M);
      SELFC_LAUNCH_CHECK("cols_to_dense_kernel");
    }
    SELFC_TRY(stp_dense<E>(ctx, i, stpbuf, pitch, feat, d, st));
    SELFC_TRY(ga_backward(ctx, ctx->ga[i], feat, gcur, gnext, grads ? grads + ga_first[i] : nullptr, wsp, ws, tp, tape, d, st));
    SELFC_TRY(dense_block_backward<E>(ctx, W, stpbuf, pitch, gnext, kStpC, gdense, scratch, grads ? grads + stp_first[i] : nullptr, d, st));
    if (i > 0) {
      copy_cols_kernel<<<cdiv(M * 16, 256), 256, 0, st>>>(gcur, kStpC, 0, gdense, pitch, 0, kStpC, M, grad_slab(ctx, d));
    } else {
      copy_cols_kernel<<<cdiv(M, 256), 256, 0, st>>>(glr, 4, 0, gdense, pitch, 0, 4, M, grad_slab(ctx, d));
    }
    SELFC_LAUNCH_CHECK("copy_cols_kernel");
  }
  // straight-through quantiser + forward-fit loss: gradient of the downscaling output
  loss_forw_kernel<<<cdiv(M, 128), 128, 0, st>>>(zs_dn + 8 * state_f, ref_l, glr, gz, lacc, kScale / n_lr, hw, M);
  SELFC_LAUNCH_CHECK("loss_forw_kernel");
  for (int blk = 7; blk >= 0; --blk)
    SELFC_TRY(invblock_backward<E>(ctx, blk, false, zs_dn + blk * state_f, gz, grads ? grads + P_INV0 + 30 * blk : nullptr, wsp, ws, d, st,
                                   keep ? &bufs_dn[blk] : nullptr, zs_dn + (blk + 1) * state_f));
  if (losses) {
    loss_finish_kernel<<<1, 1, 0, st>>>(lacc, losses, n_lr, n_hr);
    SELFC_LAUNCH_CHECK("loss_finish_kernel");
  }
  return 0;
}

// ====================================================================================================================
// Optimiser step on a FLAT gradient buffer (one NCCL all-reduce, one norm): clip_grad_norm_(params, max_norm) followed by
// torch.optim.Adam (SelfC_model.py:66-68,172-176): the parameters themselves stay separate tensors (reference layout).
// ====================================================================================================================
__global__ void __launch_bounds__(256) sqnorm_kernel(const float* __restrict__ g, long long n, float* __restrict__ out) {
  float acc = 0.f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) acc += g[i] * g[i];
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) atomicAdd(out, acc);
}

// offsets[n_tensors + 1]: element offsets of the tensors inside the flat buffers; params[t]: device pointer of tensor t
__global__ void __launch_bounds__(256) adam_kernel(float* const* __restrict__ params, const long long* __restrict__ offsets, int n_tensors,
                                                   const float* __restrict__ grad, float* __restrict__ m, float* __restrict__ v,
                                                   const float* __restrict__ sqnorm, float gscale, float max_norm, float lr, float b1,
                                                   float b2, float eps, float wd, float bc1, float bc2_sqrt) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= offsets[n_tensors]) return;
  int lo = 0, hi = n_tensors - 1;               // the tensor that holds flat element i
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (offsets[mid] <= i) lo = mid; else hi = mid - 1;
  }
  float* p = params[lo] + (i - offsets[lo]);
  float coef = gscale;
  if (max_norm > 0.f) {
    const float total = sqrtf(sqnorm[0]) * gscale;                     // norm of the (already averaged) gradient
    coef *= fminf(1.0f, max_norm / (total + 1e-6f));                    // torch.nn.utils.clip_grad_norm_
  }
  float g = grad[i] * coef;
  const float pv = *p;
  if (wd != 0.f) g = fmaf(wd, pv, g);
  const float mi = b1 * m[i] + (1.0f - b1) * g;
  const float vi = b2 * v[i] + (1.0f - b2) * g * g;
  m[i] = mi;
  v[i] = vi;
  const float denom = sqrtf(vi) / bc2_sqrt + eps;
  *p = pv - (lr / bc1) * (mi / denom);
}

// NCHW [N,51,h,w] -> planar quads (test boundary)
static int nchw51_to_quads(const float* x51, float* z, const Dims& d, cudaStream_t st) {
  SELFC_TRY(launch_nchw_slice_to_dense<float>(x51, 51, 0, z, 4, 0, 0, 3, 4, d.M(), d.hw(), st));
  for (int q = 0; q < kSQuads; ++q)
    SELFC_TRY(launch_nchw_slice_to_dense<float>(x51, 51, 3 + 4 * q, z + quad_off((size_t)d.M(), 1 + q, 0), 4, 0, 0, 4, 4, d.M(), d.hw(), st));
  return 0;
}

}  // namespace selfc

using namespace selfc;

extern "C" {

/* a13 building block: backward of D2DTInput (Subnet_constructor.py:115-133) for the dense block at `first_param`.
 * x [B*T,Cin,h,w], gy [B*T,Cout,h,w] -> gx [B*T,Cin,h,w]; gparams[10] (conv1.weight, conv1.bias, ... conv5.bias, reference
 * layouts, device, fp32) are ACCUMULATED into; any of them may be NULL.  FP32 or BF16X3 mode. */
int selfc_d2dt_backward(selfc_ctx* ctx, int first_param, const float* x, const float* gy, float* gx, float* const* gparams, int B, int T,
                        int h, int w, void* workspace, size_t workspace_bytes, void* stream) {
  Workspace ws;
  SELFC_TRY(check_run(ctx, B, T, 4 * h, 4 * w, workspace, workspace_bytes, &ws));
  SELFC_CHECK_ARG(x && gy && gx, "d2dt_backward: null pointer");
  SELFC_CHECK_ARG(ctx->mode != SELFC_MODE_BF16, "d2dt_backward: the training step runs in FP32 or BF16X3 mode");
  const DenseW* W = find_dense(ctx, first_param);
  SELFC_CHECK_ARG(W != nullptr, "d2dt_backward: parameter index %d is not the conv1.weight of a dense block", first_param);
  Dims d{B, T, h, w};
  cudaStream_t st = (cudaStream_t)stream;
  char* wsp = (char*)workspace;
  float* gbuf = reinterpret_cast<float*>(wsp + ws.params);           // M * 720 floats >= M * pitch
  float* gyd = reinterpret_cast<float*>(wsp + ws.h2);                // M * 256 floats
  float* scratch = train_scratch(ctx);                                  // dgrad weights + weight-gradient scratch (<= 0.9 MB)
  SELFC_CHECK_ARG(scratch != nullptr, "out of device memory (training scratch)");
  const int pitch = W->xpad + 4 * kGrowth;
  SELFC_CHECK_ARG(dense_bwd_scratch_floats(*W) <= kTrainScratchFloats, "d2dt_backward: scratch size");
  const int cout4 = (W->cout + 3) & ~3;
  // recompute the forward activations, then walk the block backwards
  SELFC_TRY(to_dense(gy, gyd, cout4, W->cout, cout4, d, st));
  if (ctx->mode == SELFC_MODE_BF16X3) {
    bfx2* buf = reinterpret_cast<bfx2*>(wsp + ws.stpbuf);
    SELFC_TRY(launch_nchw_to_dense<bfx2>(x, buf, pitch, dense_slab(ctx, d), 0, W->cin, W->xpad, d.M(), d.hw(), st));
    SELFC_TRY(dense_convs<bfx2>(ctx, *W, buf, pitch, d, st));
    SELFC_TRY(dense_block_backward<bfx2>(ctx, *W, buf, pitch, gyd, cout4, gbuf, scratch, gparams, d, st));
  } else {
    float* buf = reinterpret_cast<float*>(wsp + ws.stpbuf);
    SELFC_TRY(to_dense(x, buf, pitch, W->cin, W->xpad, d, st));
    SELFC_TRY(dense_convs<float>(ctx, *W, buf, pitch, d, st));
    SELFC_TRY(dense_block_backward<float>(ctx, *W, buf, pitch, gyd, cout4, gbuf, scratch, gparams, d, st));
  }
  return launch_dense_to_nchw<float>(gbuf, pitch, grad_slab(ctx, d), 0, gx, W->cin, d.M(), d.hw(), st);
}


/* a13 building block: backward of InvBlockExp (SelfC_GMM_arch_inv.py:21-33) number blk (0..7), forward (rev = 0) or reverse
 * (rev = 1) direction.  z_in [B*T,51,h,w]: the block's input; gz [B*T,51,h,w]: gradient w.r.t. the block's output on entry,
 * overwritten with the gradient w.r.t. its input; gparams[30] (F, G, H x conv1..5 weight, bias) accumulated into. */
int selfc_invblock_backward(selfc_ctx* ctx, int blk, int rev, const float* z_in, float* gz, float* const* gparams, int B, int T, int h,
                            int w, void* workspace, size_t workspace_bytes, void* stream) {
  Workspace ws;
  SELFC_TRY(check_run(ctx, B, T, 4 * h, 4 * w, workspace, workspace_bytes, &ws));
  SELFC_CHECK_ARG(z_in && gz && blk >= 0 && blk < 8, "invblock_backward: null pointer or block index");
  SELFC_CHECK_ARG(ctx->mode != SELFC_MODE_BF16, "invblock_backward: the training step runs in FP32 or BF16X3 mode");
  Dims d{B, T, h, w};
  cudaStream_t st = (cudaStream_t)stream;
  char* wsp = (char*)workspace;
  // test boundary: planar copies of the two NCHW tensors live in the STP dense buffer / feature regions
  float* zin_p = reinterpret_cast<float*>(wsp + ws.stpbuf);                  // M * 192 floats >= M * 52
  float* gz_p = zin_p + (size_t)d.M() * 4 * kZQuads;
  SELFC_TRY(nchw51_to_quads(z_in, zin_p, d, st));
  SELFC_TRY(nchw51_to_quads(gz, gz_p, d, st));
  if (ctx->mode == SELFC_MODE_BF16X3) SELFC_TRY(invblock_backward<bfx2>(ctx, blk, rev != 0, zin_p, gz_p, gparams, wsp, ws, d, st));
  else SELFC_TRY(invblock_backward<float>(ctx, blk, rev != 0, zin_p, gz_p, gparams, wsp, ws, d, st));
  return launch_export_down(gz_p, gz, nullptr, nullptr, d.M(), d.hw(), st);
}

size_t selfc_train_tape_bytes(int B, int T, int h, int w) {
  if (B < 1 || T < 1 || h < 1 || w < 1) return 0;
  return make_tape(B, T, h, w).total;
}

/* a13 building block: backward of the GMM head (tail_gmm, :336-344) + soft-GMM sampler (:383-394).  feat [B*T,64,h,w]: the STP
 * feature; gv [B*T,48,h,w]: gradient w.r.t. the sampled HF latents; eps as in selfc_up (NULL: Philox stream seed/offset);
 * gfeat [B*T,64,h,w] out; gparams[6]: tail_gmm.{1,3,5}.{weight,bias} gradients, accumulated into. */
int selfc_head_sampler_backward(selfc_ctx* ctx, const float* feat, const float* gv, const float* eps, uint64_t seed, uint64_t offset,
                                float* gfeat, float* const* gparams, int B, int T, int h, int w, void* workspace, size_t workspace_bytes,
                                void* tape, size_t tape_bytes, void* stream) {
  Workspace ws;
  SELFC_TRY(check_run(ctx, B, T, 4 * h, 4 * w, workspace, workspace_bytes, &ws));
  SELFC_CHECK_ARG(feat && gv && gfeat && tape, "head_sampler_backward: null pointer");
  SELFC_CHECK_ARG(ctx->mode != SELFC_MODE_BF16, "head_sampler_backward: the training step runs in FP32 or BF16X3 mode");
  const Tape tl = make_tape(B, T, h, w);
  SELFC_CHECK_ARG(tape_bytes >= tl.total && aligned16(tape), "head_sampler_backward: tape too small (%zu < %zu)", tape_bytes, tl.total);
  Dims d{B, T, h, w};
  cudaStream_t st = (cudaStream_t)stream;
  char* tp = (char*)tape;
  float* featd = reinterpret_cast<float*>(tp + tl.ga) + (size_t)6 * d.M() * 64;
  float* gz = reinterpret_cast<float*>(tp + tl.gz);
  float* gfd = reinterpret_cast<float*>(tp + tl.ga);               // slot 0 as the output scratch
  SELFC_TRY(launch_nchw_to_dense<float>(feat, featd, 64, 0, 0, 64, 64, d.M(), d.hw(), st));
  for (int q = 0; q < kSQuads; ++q)
    SELFC_TRY(launch_nchw_slice_to_dense<float>(gv, kHF, 4 * q, gz + quad_off((size_t)d.M(), 1 + q, 0), 4, 0, 0, 4, 4, d.M(), d.hw(), st));
  SELFC_TRY(head_sampler_backward(ctx, featd, eps, seed, offset, gz, gfd, gparams, (char*)workspace, ws, tp, tl, d, st));
  return launch_dense_to_nchw<float>(gfd, 64, 0, 0, gfeat, 64, d.M(), d.hw(), st);
}

/* a13 building block: backward of GlobalAgg (:265-285) whose fc.weight is parameter `first_param`.  x, gout [B*T,64,h,w] -> gx;
 * gparams[8] (fc.weight, fc.bias, proj1.weight, proj1.bias, proj2.weight, proj2.bias, proj3.weight, proj3.bias) accumulated into. */
int selfc_global_agg_backward(selfc_ctx* ctx, int first_param, const float* x, const float* gout, float* gx, float* const* gparams, int B,
                              int T, int h, int w, void* workspace, size_t workspace_bytes, void* tape, size_t tape_bytes, void* stream) {
  Workspace ws;
  SELFC_TRY(check_run(ctx, B, T, 4 * h, 4 * w, workspace, workspace_bytes, &ws));
  SELFC_CHECK_ARG(x && gout && gx && tape, "global_agg_backward: null pointer");
  SELFC_CHECK_ARG(ctx->mode != SELFC_MODE_BF16 && T <= 16, "global_agg_backward: FP32 / BF16X3 mode and T <= 16 only");
  const GaW* g = find_ga(ctx, first_param);
  SELFC_CHECK_ARG(g != nullptr, "global_agg_backward: parameter index %d is not the fc.weight of a GlobalAgg", first_param);
  const Tape tl = make_tape(B, T, h, w);
  SELFC_CHECK_ARG(tape_bytes >= tl.total && aligned16(tape), "global_agg_backward: tape too small (%zu < %zu)", tape_bytes, tl.total);
  Dims d{B, T, h, w};
  cudaStream_t st = (cudaStream_t)stream;
  char* tp = (char*)tape;
  float* xd = reinterpret_cast<float*>(tp + tl.ga);
  float* gd = xd + (size_t)d.M() * 64;
  float* gxd = gd + (size_t)d.M() * 64;
  SELFC_TRY(launch_nchw_to_dense<float>(x, xd, 64, 0, 0, 64, 64, d.M(), d.hw(), st));
  SELFC_TRY(launch_nchw_to_dense<float>(gout, gd, 64, 0, 0, 64, 64, d.M(), d.hw(), st));
  SELFC_TRY(ga_backward(ctx, *g, xd, gd, gxd, gparams, (char*)workspace, ws, tp, tl, d, st));
  return launch_dense_to_nchw<float>(gxd, 64, 0, 0, gx, 64, d.M(), d.hw(), st);
}

/* a13: forward + backward of one training step (models/SelfC_model.py:148-170 with the training YAML's losses: l2 forward fit
 * against ref_l, Charbonnier reconstruction, x 144*144*3; Quantization with its straight-through gradient).  hr [B*T,3,H,W],
 * ref_l [B*T,3,H/4,W/4]; eps as in selfc_up.  grads[354]: device fp32 buffers in the reference parameter layouts, ACCUMULATED
 * into (zero them for a plain step); losses: 3 device floats (total, l_forw_fit, l_back_rec).  FP32 or BF16X3 mode, T <= 16. */
int selfc_train_grads(selfc_ctx* ctx, const float* hr, const float* ref_l, const float* eps, uint64_t seed, uint64_t offset,
                      float* const* grads, int n_grads, float* losses, int B, int T, int H, int W, void* workspace, size_t workspace_bytes,
                      void* tape, size_t tape_bytes, void* stream) {
  Workspace ws;
  SELFC_TRY(check_run(ctx, B, T, H, W, workspace, workspace_bytes, &ws));
  SELFC_CHECK_ARG(hr && ref_l && tape && aligned16(hr), "train_grads: null or misaligned pointer");
  SELFC_CHECK_ARG(grads == nullptr || n_grads == SELFC_NUM_PARAMS, "train_grads: expected %d gradient buffers", SELFC_NUM_PARAMS);
  SELFC_CHECK_ARG(ctx->mode != SELFC_MODE_BF16 && T <= 16, "train_grads: FP32 / BF16X3 mode and T <= 16 only");
  const Tape tl = make_tape(B, T, H / 4, W / 4);
  SELFC_CHECK_ARG(tape_bytes >= tl.total && aligned16(tape), "train_grads: tape too small (%zu < %zu)", tape_bytes, tl.total);
  Dims d{B, T, H / 4, W / 4};
  if (ctx->mode == SELFC_MODE_BF16X3)
    return train_grads<bfx2>(ctx, hr, ref_l, eps, seed, offset, grads, losses, d, (char*)workspace, ws, (char*)tape, tl, (cudaStream_t)stream);
  return train_grads<float>(ctx, hr, ref_l, eps, seed, offset, grads, losses, d, (char*)workspace, ws, (char*)tape, tl, (cudaStream_t)stream);
}

/* a13 optimiser step (SelfC_model.py:66-68,172-176): gradient clipping by global norm + Adam, on a flat gradient buffer.
 * params[n_tensors]: DEVICE array of device pointers to the parameter tensors; offsets[n_tensors+1]: DEVICE array of element
 * offsets of those tensors inside grad / m / v (flat fp32 buffers of offsets[n_tensors] elements, given as total);
 * gscale multiplies the gradient first (1/world after a SUM all-reduce); max_norm <= 0 disables clipping; step >= 1.
 * sqnorm_scratch: one device float. */
int selfc_adam_step(float* const* params, const long long* offsets, int n_tensors, long long total, const float* grad, float* m,
                    float* v, float* sqnorm_scratch, float gscale, float max_norm, float lr, float beta1, float beta2, float eps,
                    float weight_decay, int step, void* stream) {
  SELFC_CHECK_ARG(params && offsets && grad && m && v && sqnorm_scratch && n_tensors >= 1 && total >= 1 && step >= 1, "adam_step: bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  SELFC_CUDA(cudaMemsetAsync(sqnorm_scratch, 0, sizeof(float), st));
  if (max_norm > 0.f) {
    sqnorm_kernel<<<296, 256, 0, st>>>(grad, total, sqnorm_scratch);
    SELFC_LAUNCH_CHECK("sqnorm_kernel");
  }
  const float bc1 = 1.0f - powf(beta1, (float)step);
  const float bc2_sqrt = sqrtf(1.0f - powf(beta2, (float)step));
  adam_kernel<<<cdiv(total, 256), 256, 0, st>>>(params, offsets, n_tensors, grad, m, v, sqnorm_scratch, gscale, max_norm, lr, beta1, beta2, eps,
                                               weight_decay, bc1, bc2_sqrt);
  SELFC_LAUNCH_CHECK("adam_kernel");
  return 0;
}
}  // extern "C"
