// Backward kernels of the training step (models/SelfC_model.py:148-183) -- FP32 mode, pixel-major dense
// buffers, fp32-FMA arithmetic.  First building block: the backward pass of one D2DTInput dense block
// (Subnet_constructor.py:115-133), parity-tested against autograd of the oracle.
//
//   forward   x_k = lrelu(conv_k([X, x_1..x_{k-1}])), k = 1..4 (1,3,3);   y = conv5([X, x_1..x_4]) (3,1,1), no activation
//   backward  with gbuf = gradient w.r.t. the dense buffer [X | x1 | x2 | x3 | x4] (zero-initialised):
//     conv5:  dW5 += g_y (x) in5,   gbuf[0:cin5) += conv5^T(g_y)
//     k=4..1: g_k = gbuf[slot_k] * lrelu'(x_k);   dW_k += g_k (x) in_k;   gbuf[0:cin_k) += conv_k^T(g_k)
//     g_X = gbuf[0:cin)
//   conv^T (dgrad) is the SAME implicit-GEMM kernel as the forward conv (conv_simt.cu) on the tap-flipped, transposed
//   weights, accumulating into gbuf (EPI_ACCUM); wgrad is a pixel reduction (wgrad_kernel below).
#include <string.h>

#include "net_ctx.h"

namespace selfc {

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
__global__ void lrelu_bwd_kernel(float* __restrict__ g, const float* __restrict__ y, int pitch, int off, long long M) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long m = idx >> 3;
  const int c = (int)(idx & 7) * 4;
  if (m >= M) return;
  float4 gv = load4(g + m * pitch + off + c);
  const float4 yv = load4(y + m * pitch + off + c);
  gv.x *= yv.x > 0.f ? 1.f : 0.2f;
  gv.y *= yv.y > 0.f ? 1.f : 0.2f;
  gv.z *= yv.z > 0.f ? 1.f : 0.2f;
  gv.w *= yv.w > 0.f ? 1.f : 0.2f;
  store4(g + m * pitch + off + c, gv);
}

// wgrad: dw[tap][c][n] += sum_m in[m + shift(tap)][c] * g[m][n]  (+ the bias gradient in pseudo-tap `taps`: db[n] += sum_m g[m][n]).
// grid (taps + 1, ceil(cin/32), splits); block 256 = 32 (c) x 8 (n groups of 4); each CTA reduces its share of the pixels through
// shared-memory tiles of 32 pixels and adds its partial 32x32 block with atomics (fp32 scratch in buffer-channel order).
__global__ void __launch_bounds__(256) wgrad_kernel(const float* __restrict__ in, int in_pitch, int cin, const float* __restrict__ g,
                                                    int g_pitch, int g_off, int cout, float* __restrict__ dw, int np, int taps,
                                                    int tap_mode, int BT, int Tn, int h, int w_) {
  __shared__ float As[32][33];     // [pixel][c]
  __shared__ float Gs[32][33];     // [pixel][n]
  const int tap = blockIdx.x % (taps + 1);
  const int n0 = (blockIdx.x / (taps + 1)) * 32;          // output-channel tile
  const int c0 = blockIdx.y * 32;
  const bool bias_pass = tap == taps;
  if (bias_pass && blockIdx.y != 0) return;
  const long long hw = (long long)h * w_;
  const long long M = (long long)BT * hw;
  const int tid = threadIdx.x;
  const int tc = tid & 31, tn = (tid >> 5) * 4;
  int dy = 0, dx = 0, dt = 0;
  if (!bias_pass) {
    if (tap_mode == TAP_SPATIAL) { dy = tap / 3 - 1; dx = tap % 3 - 1; }
    else if (tap_mode == TAP_TEMPORAL) dt = tap - 1;
  }
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  const long long per = (M + gridDim.z - 1) / gridDim.z;
  const long long m_begin = (long long)blockIdx.z * per;
  const long long m_end = m_begin + per < M ? m_begin + per : M;
  for (long long mb = m_begin; mb < m_end; mb += 32) {
    // stage 32 pixels: thread (p = tid/8, q = tid%8) loads 4 channels of A and 4 of G
    const int p = tid >> 3, q4 = (tid & 7) * 4;
    const long long m = mb + p;
    float4 av = make_float4(0.f, 0.f, 0.f, 0.f), gv = make_float4(0.f, 0.f, 0.f, 0.f);
    if (m < m_end) {
      if (n0 + q4 < cout) {
        const float* gp = g + m * g_pitch + g_off + n0 + q4;
        gv.x = gp[0];
        if (n0 + q4 + 1 < cout) gv.y = gp[1];
        if (n0 + q4 + 2 < cout) gv.z = gp[2];
        if (n0 + q4 + 3 < cout) gv.w = gp[3];
      }
      if (bias_pass) {
        av = make_float4(1.f, 1.f, 1.f, 1.f);
      } else if (c0 + q4 < cin) {
        const long long n = m / hw, pix = m - n * hw;
        const int y = (int)(pix / w_), x = (int)(pix - (long long)y * w_);
        const int t = (int)(n % Tn);
        const bool ok = (unsigned)(y + dy) < (unsigned)h && (unsigned)(x + dx) < (unsigned)w_ && (unsigned)(t + dt) < (unsigned)Tn;
        if (ok) av = load4(in + (m + (long long)dt * hw + dy * w_ + dx) * in_pitch + c0 + q4);
      }
    }
    As[p][q4] = av.x; As[p][q4 + 1] = av.y; As[p][q4 + 2] = av.z; As[p][q4 + 3] = av.w;
    Gs[p][q4] = gv.x; Gs[p][q4 + 1] = gv.y; Gs[p][q4 + 2] = gv.z; Gs[p][q4 + 3] = gv.w;
    __syncthreads();
#pragma unroll 8
    for (int pp = 0; pp < 32; ++pp) {
      const float a = As[pp][tc];
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[j] = fmaf(a, Gs[pp][tn + j], acc[j]);
    }
    __syncthreads();
  }
  if (bias_pass) {
    if (tc == 0)
      for (int j = 0; j < 4; ++j)
        if (n0 + tn + j < cout) atomicAdd(dw + (long long)taps * cin * np + n0 + tn + j, acc[j]);
  } else if (c0 + tc < cin) {
    for (int j = 0; j < 4; ++j)
      if (n0 + tn + j < cout) atomicAdd(dw + ((long long)tap * cin + c0 + tc) * np + n0 + tn + j, acc[j]);
  }
}

// scratch [taps][cin_buf][np] (+ [np] bias) in buffer-channel order -> reference layouts dW [cout][cin_ref][taps], db [cout] (accumulating)
__global__ void wgrad_unpack_kernel(const float* __restrict__ dw, float* __restrict__ gw, float* __restrict__ gb, int cout, int cin_ref,
                                    int taps, int cin_buf, int xreal, int xpad, int np) {
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

int dense_block_backward(const selfc_ctx* ctx, const DenseW& W, const float* buf, int pitch, const float* gy, int gy_pitch, float* gbuf,
                         float* scratch, float* const* gparams, const Dims& d, cudaStream_t st) {
  SELFC_CHECK_ARG(ctx->mode == SELFC_MODE_FP32, "the training step runs in FP32 mode");
  const long long M = d.M();
  if (M == 0) return 0;
  SELFC_CUDA(cudaMemsetAsync(gbuf, 0, (size_t)M * pitch * sizeof(float), st));
  static float* zero_bias_dev[64] = {};       // dgrad has no bias term
  int dev = 0;
  cudaGetDevice(&dev);
  SELFC_CHECK_ARG(dev >= 0 && dev < 64, "device index");
  if (!zero_bias_dev[dev]) {
    SELFC_CUDA(cudaMalloc(&zero_bias_dev[dev], 256 * sizeof(float)));
    SELFC_CUDA(cudaMemset(zero_bias_dev[dev], 0, 256 * sizeof(float)));
  }
  float* zero_bias = zero_bias_dev[dev];
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
      lrelu_bwd_kernel<<<cdiv(M * 8, 256), 256, 0, st>>>(gbuf, buf, pitch, slot, M);
      SELFC_LAUNCH_CHECK("lrelu_bwd_kernel");
    }
    float* wd = scratch;
    float* dw = scratch + (size_t)taps * cout4 * npd;
    // weight / bias gradients
    if (gparams != nullptr && gparams[2 * k] != nullptr) {
      const size_t dw_floats = (size_t)(taps * cin + 1) * W.np[k];
      SELFC_CUDA(cudaMemsetAsync(dw, 0, dw_floats * sizeof(float), st));
      int splits = (int)(M / 2048);
      if (splits < 1) splits = 1;
      if (splits > 64) splits = 64;
      dim3 grid((taps + 1) * cdiv(cout, 32), cdiv(cin, 32), splits);
      wgrad_kernel<<<grid, 256, 0, st>>>(buf, pitch, cin, g, g_pitch, g_off, cout, dw, W.np[k], taps, tap_mode, d.B * d.T, d.T, d.h, d.w);
      SELFC_LAUNCH_CHECK("wgrad_kernel");
      const int cin_ref = W.cin + kGrowth * k;
      const long long total = (long long)cout * cin_ref * taps;
      wgrad_unpack_kernel<<<cdiv(total > cout ? total : cout, 256), 256, 0, st>>>(dw, gparams[2 * k], gparams[2 * k + 1], cout, cin_ref, taps, cin,
                                                                                W.cin, W.xpad, W.np[k]);
      SELFC_LAUNCH_CHECK("wgrad_unpack_kernel");
    }
    // input gradient: the forward kernel on the flipped / transposed weights, accumulated into gbuf[0:cin)
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

}  // namespace selfc

using namespace selfc;

extern "C" {

/* a13 building block: backward of D2DTInput (Subnet_constructor.py:115-133) for the dense block at `first_param`.
 * x [B*T,Cin,h,w], gy [B*T,Cout,h,w] -> gx [B*T,Cin,h,w]; gparams[10] (conv1.weight, conv1.bias, ... conv5.bias, reference
 * layouts, device, fp32) are ACCUMULATED into; any of them may be NULL.  FP32 mode only. */
int selfc_d2dt_backward(selfc_ctx* ctx, int first_param, const float* x, const float* gy, float* gx, float* const* gparams, int B, int T,
                        int h, int w, void* workspace, size_t workspace_bytes, void* stream) {
  Workspace ws;
  SELFC_TRY(check_run(ctx, B, T, 4 * h, 4 * w, workspace, workspace_bytes, &ws));
  SELFC_CHECK_ARG(x && gy && gx, "d2dt_backward: null pointer");
  SELFC_CHECK_ARG(ctx->mode == SELFC_MODE_FP32, "d2dt_backward: the training step runs in FP32 mode");
  const DenseW* W = find_dense(ctx, first_param);
  SELFC_CHECK_ARG(W != nullptr, "d2dt_backward: parameter index %d is not the conv1.weight of a dense block", first_param);
  Dims d{B, T, h, w};
  cudaStream_t st = (cudaStream_t)stream;
  char* wsp = (char*)workspace;
  float* buf = reinterpret_cast<float*>(wsp + ws.stpbuf);
  float* gbuf = reinterpret_cast<float*>(wsp + ws.params);           // M * 720 floats >= M * pitch
  float* gyd = reinterpret_cast<float*>(wsp + ws.h2);                // M * 256 floats
  static float* scratch_dev[64] = {};                                // dgrad weights + weight-gradient scratch (<= 0.9 MB)
  int dev = 0;
  cudaGetDevice(&dev);
  SELFC_CHECK_ARG(dev >= 0 && dev < 64, "device index");
  if (!scratch_dev[dev]) SELFC_CUDA(cudaMalloc(&scratch_dev[dev], (size_t)(9 * 64 * 192 + (9 * 192 + 1) * 64 + 256) * sizeof(float)));
  float* scratch = scratch_dev[dev];
  const int pitch = W->xpad + 4 * kGrowth;
  SELFC_CHECK_ARG(dense_bwd_scratch_floats(*W) <= (size_t)(9 * 64 * 192 + (9 * 192 + 1) * 64 + 256), "d2dt_backward: scratch size");
  const int cout4 = (W->cout + 3) & ~3;
  // recompute the forward activations, then walk the block backwards
  SELFC_TRY(to_dense(x, buf, pitch, W->cin, W->xpad, d, st));
  SELFC_TRY(dense_convs_f32(ctx, *W, buf, pitch, d, st));
  SELFC_TRY(to_dense(gy, gyd, cout4, W->cout, cout4, d, st));
  SELFC_TRY(dense_block_backward(ctx, *W, buf, pitch, gyd, cout4, gbuf, scratch, gparams, d, st));
  return launch_dense_to_nchw<float>(gbuf, pitch, 0, 0, gx, W->cin, d.M(), d.hw(), st);
}

}  // extern "C"
