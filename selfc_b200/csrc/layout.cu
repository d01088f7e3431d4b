// Memory-bound layout kernels of the rescaling path: FrequencyAnalyzer forward/reverse, the LR split +
// 8-bit quantisation, and the NCHW <-> pixel-major ("dense buffer") conversions at the boundary.
//
// Reference behaviour restated (not copied): models/modules/SelfC_GMM_arch_inv.py:46-82 (FrequencyAnalyzer,
// PixelUnshuffle), models/modules/Quantization.py:4-17, models/SelfC_model.py:217-222.
//
// Internal layout: latent state z as planar quads [13][M][4] fp32, M = B*T*h*w (common.cuh); quad 0 = LR part x1,
// quads 1..12 = HF part x2 (in the FORWARD channel order (sy*4+sx)*3+c while going down, and whatever the
// couplings produce while going up -- the reverse FrequencyAnalyzer reads it as c*16+sy*4+sx).
#include "common.cuh"
#include "kernels.h"

namespace selfc {

// ------------------------------------------------------------------------------------------------------
// FrequencyAnalyzer forward: one thread per LR pixel, 12 x 16-byte loads (4 rows x 3 colours).
// OUT_NCHW: write [N,51,h,w] (standalone component); else write z [M][52] (+ optional T copy of the 48 HF
// channels into the F dense buffer, channel offset 0).
// ------------------------------------------------------------------------------------------------------
// IN_U8: x is a decoded 8-bit frame stack [N][H][W][3] in cv2 order (B,G,R); value = float(byte)/255.0f, the
// conversion read_img1 + the dataset's BGR->RGB / HWC->CHW make on the CPU (data/util.py:103-115,
// data/LQGTVID_dataset.py:150-154), done in the load instead.
template <bool OUT_NCHW, typename T, bool IN_U8 = false>
__global__ void __launch_bounds__(256) fa_fwd_kernel(const void* __restrict__ xin, float* __restrict__ out,
                                                     T* __restrict__ fbuf, int fpitch, long long fslabM, int N, int h, int w) {
  const long long M = (long long)N * h * w;
  const long long m = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= M) return;
  const int j = (int)(m % w);
  const int i = (int)((m / w) % h);
  const int n = (int)(m / ((long long)w * h));
  const int W = 4 * w, H = 4 * h;
  float v[3][16];
  if (IN_U8) {
    // 4 rows x 12 bytes (4 pixels x BGR); 12-byte aligned because W % 4 == 0
    const uint8_t* src = reinterpret_cast<const uint8_t*>(xin) + (((long long)n * H + 4 * i) * W + 4 * j) * 3;
#pragma unroll
    for (int sy = 0; sy < 4; ++sy) {
      const uint32_t* r = reinterpret_cast<const uint32_t*>(src + (long long)sy * W * 3);
      uint32_t wds[3] = {__ldg(r), __ldg(r + 1), __ldg(r + 2)};
#pragma unroll
      for (int b = 0; b < 12; ++b) {
        const uint32_t byte = (wds[b >> 2] >> (8 * (b & 3))) & 0xffu;
        v[2 - b % 3][sy * 4 + b / 3] = (float)byte / 255.0f;
      }
    }
  } else {
    const float* x = reinterpret_cast<const float*>(xin);
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float* src = x + (((long long)n * 3 + c) * H + 4 * i) * W + 4 * j;
#pragma unroll
      for (int sy = 0; sy < 4; ++sy) {
        float4 r = __ldg(reinterpret_cast<const float4*>(src + (long long)sy * W));
        v[c][sy * 4 + 0] = r.x; v[c][sy * 4 + 1] = r.y; v[c][sy * 4 + 2] = r.z; v[c][sy * 4 + 3] = r.w;
      }
    }
  }
  float lf[3];
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    // sequential row-major accumulation then /16: the order adaptive_avg_pool2d uses (bit-exact, oracle fa_forward)
    float acc = 0.f;
#pragma unroll
    for (int q = 0; q < 16; ++q) acc = acc + v[c][q];
    lf[c] = acc / 16.0f;
  }
  if (OUT_NCHW) {
    const long long hw = (long long)h * w;
    float* o = out + (long long)n * 51 * hw + (long long)i * w + j;
#pragma unroll
    for (int c = 0; c < 3; ++c) o[c * hw] = lf[c];
#pragma unroll
    for (int q = 0; q < 16; ++q)
#pragma unroll
      for (int c = 0; c < 3; ++c) o[(3 + q * 3 + c) * hw] = v[c][q] - lf[c];
  } else {
    float row[4 * kZQuads];
    row[0] = lf[0]; row[1] = lf[1]; row[2] = lf[2]; row[3] = 0.f;
#pragma unroll
    for (int q = 0; q < 16; ++q)
#pragma unroll
      for (int c = 0; c < 3; ++c) row[4 + q * 3 + c] = v[c][q] - lf[c];
#pragma unroll
    for (int k = 0; k < kZQuads; ++k)
      store4(out + quad_off(M, k, m), make_float4(row[4 * k], row[4 * k + 1], row[4 * k + 2], row[4 * k + 3]));
    if (fbuf != nullptr) {
#pragma unroll
      for (int k = 0; k < kHF; k += 4)
        store4(fbuf + dense_off(m, k, fpitch, fslabM), make_float4(row[4 + k], row[4 + k + 1], row[4 + k + 2], row[4 + k + 3]));
    }
  }
}

// ------------------------------------------------------------------------------------------------------
// FrequencyAnalyzer reverse: y[n,c,4i+sy,4j+sx] = lf[c] + hf[c*16+sy*4+sx]  (nn.PixelShuffle order).
// ------------------------------------------------------------------------------------------------------
// OUT_U8: write the frames as 8-bit [N][H][W][3] in cv2 order (B,G,R): round-half-even of clamp(v,0,1)*255, the
// conversion tensor2img makes on the CPU before save_img (utils/util.py:104-133,181-182).
__device__ __forceinline__ uint32_t img_code(float v) {
  v = fminf(fmaxf(v, 0.f), 1.f);
  return (uint32_t)rintf(v * 255.0f);
}

template <bool IN_NCHW, bool OUT_U8 = false>
__global__ void __launch_bounds__(256) fa_rev_kernel(const float* __restrict__ z, void* __restrict__ yout, int N, int h, int w) {
  const long long M = (long long)N * h * w;
  const long long m = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= M) return;
  const int j = (int)(m % w);
  const int i = (int)((m / w) % h);
  const int n = (int)(m / ((long long)w * h));
  const int W = 4 * w, H = 4 * h;
  float lf[3], hf[kHF];
  if (IN_NCHW) {
    const long long hw = (long long)h * w;
    const float* src = z + (long long)n * 51 * hw + (long long)i * w + j;
#pragma unroll
    for (int c = 0; c < 3; ++c) lf[c] = __ldg(src + c * hw);
#pragma unroll
    for (int k = 0; k < kHF; ++k) hf[k] = __ldg(src + (3 + k) * hw);
  } else {
    float4 a = load4(z + quad_off(M, 0, m));
    lf[0] = a.x; lf[1] = a.y; lf[2] = a.z;
#pragma unroll
    for (int k = 0; k < kHF; k += 4) {
      float4 r = load4(z + quad_off(M, 1 + k / 4, m));
      hf[k] = r.x; hf[k + 1] = r.y; hf[k + 2] = r.z; hf[k + 3] = r.w;
    }
  }
  if (OUT_U8) {
    uint8_t* dst = reinterpret_cast<uint8_t*>(yout) + (((long long)n * H + 4 * i) * W + 4 * j) * 3;
#pragma unroll
    for (int sy = 0; sy < 4; ++sy) {
      uint32_t wds[3] = {0u, 0u, 0u};
#pragma unroll
      for (int b = 0; b < 12; ++b) {
        const int c = 2 - b % 3, sx = b / 3;
        wds[b >> 2] |= img_code(lf[c] + hf[c * 16 + sy * 4 + sx]) << (8 * (b & 3));
      }
      uint32_t* r = reinterpret_cast<uint32_t*>(dst + (long long)sy * W * 3);
      r[0] = wds[0]; r[1] = wds[1]; r[2] = wds[2];
    }
    return;
  }
  float* y = reinterpret_cast<float*>(yout);
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    float* dst = y + (((long long)n * 3 + c) * H + 4 * i) * W + 4 * j;
#pragma unroll
    for (int sy = 0; sy < 4; ++sy) {
      float4 r = make_float4(lf[c] + hf[c * 16 + sy * 4 + 0], lf[c] + hf[c * 16 + sy * 4 + 1],
                             lf[c] + hf[c * 16 + sy * 4 + 2], lf[c] + hf[c * 16 + sy * 4 + 3]);
      *reinterpret_cast<float4*>(dst + (long long)sy * W) = r;
    }
  }
}

// ------------------------------------------------------------------------------------------------------
// Stand-alone 8-bit frame conversions (any H, W): [N][H][W][3] BGR bytes <-> [N][3][H][W] RGB fp32.  One thread per
// pixel; the LR side of the path uses them (LR is 1/16 of the HR pixels), the HR side is fused into the
// FrequencyAnalyzer kernels above.
// ------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) frames_from_u8_kernel(const uint8_t* __restrict__ img, float* __restrict__ x, long long N, long long HW) {
  const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= N * HW) return;
  const long long n = p / HW, q = p - n * HW;
  const uint8_t* s = img + p * 3;
  float* d = x + n * 3 * HW + q;
  d[0] = (float)s[2] / 255.0f;
  d[HW] = (float)s[1] / 255.0f;
  d[2 * HW] = (float)s[0] / 255.0f;
}

__global__ void __launch_bounds__(256) frames_to_u8_kernel(const float* __restrict__ x, uint8_t* __restrict__ img, long long N, long long HW) {
  const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= N * HW) return;
  const long long n = p / HW, q = p - n * HW;
  const float* s = x + n * 3 * HW + q;
  uint8_t* d = img + p * 3;
  d[0] = (uint8_t)img_code(s[2 * HW]);
  d[1] = (uint8_t)img_code(s[HW]);
  d[2] = (uint8_t)img_code(s[0]);
}

// ------------------------------------------------------------------------------------------------------
// 2x variants.  FrequencyAnalyzer(k=2) of the compression model's rescaler half
// (SelfC_Codec_arch_inv.py:78-98): [N,3,H,W] <-> [N,15,H/2,W/2], same structure as the 4x one (2x2 box mean + residual;
// forward HF channel = (sy*2+sx)*3+c, reverse = c*4+sy*2+sx).  HaarDownsampling of `model: SelfC` / IRN
// (SelfC_arch_inv.py:44-84): grouped 2x2 stride-2 conv with +-1 weights, /4, channel k*C+c; the reverse is the transposed
// conv with the same weights and no scaling.  One thread per output LR pixel (x channel for Haar).
// ------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) fa2_fwd_kernel(const float* __restrict__ x, float* __restrict__ out, int N, int h, int w) {
  const long long M = (long long)N * h * w;
  const long long m = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= M) return;
  const int j = (int)(m % w), i = (int)((m / w) % h), n = (int)(m / ((long long)w * h));
  const int W = 2 * w, H = 2 * h;
  const long long hw = (long long)h * w;
  float* o = out + (long long)n * 15 * hw + (long long)i * w + j;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const float* src = x + (((long long)n * 3 + c) * H + 2 * i) * W + 2 * j;
    const float2 r0 = __ldg(reinterpret_cast<const float2*>(src));
    const float2 r1 = __ldg(reinterpret_cast<const float2*>(src + W));
    const float lf = (((r0.x + r0.y) + r1.x) + r1.y) / 4.0f;       // row-major accumulation, as adaptive_avg_pool2d
    o[c * hw] = lf;
    o[(3 + 0 * 3 + c) * hw] = r0.x - lf;
    o[(3 + 1 * 3 + c) * hw] = r0.y - lf;
    o[(3 + 2 * 3 + c) * hw] = r1.x - lf;
    o[(3 + 3 * 3 + c) * hw] = r1.y - lf;
  }
}

__global__ void __launch_bounds__(256) fa2_rev_kernel(const float* __restrict__ z, float* __restrict__ y, int N, int h, int w) {
  const long long M = (long long)N * h * w;
  const long long m = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= M) return;
  const int j = (int)(m % w), i = (int)((m / w) % h), n = (int)(m / ((long long)w * h));
  const int W = 2 * w, H = 2 * h;
  const long long hw = (long long)h * w;
  const float* src = z + (long long)n * 15 * hw + (long long)i * w + j;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const float lf = __ldg(src + c * hw);
    float* dst = y + (((long long)n * 3 + c) * H + 2 * i) * W + 2 * j;
    *reinterpret_cast<float2*>(dst) = make_float2(lf + __ldg(src + (3 + c * 4 + 0) * hw), lf + __ldg(src + (3 + c * 4 + 1) * hw));
    *reinterpret_cast<float2*>(dst + W) = make_float2(lf + __ldg(src + (3 + c * 4 + 2) * hw), lf + __ldg(src + (3 + c * 4 + 3) * hw));
  }
}

// Haar weights w_k[dy][dx]: k=0 all +1; k=1 -1 on dx=1; k=2 -1 on dy=1; k=3 -1 on the anti-diagonal
__global__ void __launch_bounds__(256) haar_fwd_kernel(const float* __restrict__ x, float* __restrict__ out, int N, int C, int h, int w) {
  const long long M = (long long)N * C * h * w;
  const long long m = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= M) return;
  const int j = (int)(m % w), i = (int)((m / w) % h);
  const int c = (int)((m / ((long long)w * h)) % C), n = (int)(m / ((long long)w * h * C));
  const int W = 2 * w;
  const long long hw = (long long)h * w;
  const float* src = x + (((long long)n * C + c) * 2 * h + 2 * i) * W + 2 * j;
  const float2 r0 = __ldg(reinterpret_cast<const float2*>(src));
  const float2 r1 = __ldg(reinterpret_cast<const float2*>(src + W));
  const float a = r0.x, b = r0.y, cc = r1.x, d = r1.y;
  float* o = out + ((long long)n * 4 * C + c) * hw + (long long)i * w + j;
  o[0] = (((a + b) + cc) + d) / 4.0f;
  o[(long long)C * hw] = (((a - b) + cc) - d) / 4.0f;
  o[2LL * C * hw] = (((a + b) - cc) - d) / 4.0f;
  o[3LL * C * hw] = (((a - b) - cc) + d) / 4.0f;
}

__global__ void __launch_bounds__(256) haar_rev_kernel(const float* __restrict__ z, float* __restrict__ y, int N, int C, int h, int w) {
  const long long M = (long long)N * C * h * w;
  const long long m = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= M) return;
  const int j = (int)(m % w), i = (int)((m / w) % h);
  const int c = (int)((m / ((long long)w * h)) % C), n = (int)(m / ((long long)w * h * C));
  const int W = 2 * w;
  const long long hw = (long long)h * w;
  const float* src = z + ((long long)n * 4 * C + c) * hw + (long long)i * w + j;
  const float z0 = __ldg(src), z1 = __ldg(src + (long long)C * hw), z2 = __ldg(src + 2LL * C * hw), z3 = __ldg(src + 3LL * C * hw);
  float* dst = y + (((long long)n * C + c) * 2 * h + 2 * i) * W + 2 * j;
  *reinterpret_cast<float2*>(dst) = make_float2(((z0 + z1) + z2) + z3, ((z0 - z1) + z2) - z3);
  *reinterpret_cast<float2*>(dst + W) = make_float2(((z0 + z1) - z2) - z3, ((z0 - z1) - z2) + z3);
}

int launch_fa2(const float* in, float* out, bool rev, int N, int h, int w, cudaStream_t st) {
  const long long M = (long long)N * h * w;
  if (M == 0) return 0;
  if (rev) fa2_rev_kernel<<<cdiv(M, 256), 256, 0, st>>>(in, out, N, h, w);
  else fa2_fwd_kernel<<<cdiv(M, 256), 256, 0, st>>>(in, out, N, h, w);
  SELFC_LAUNCH_CHECK("fa2_kernel");
  return 0;
}

int launch_haar(const float* in, float* out, bool rev, int N, int C, int h, int w, cudaStream_t st) {
  const long long M = (long long)N * C * h * w;
  if (M == 0) return 0;
  if (rev) haar_rev_kernel<<<cdiv(M, 256), 256, 0, st>>>(in, out, N, C, h, w);
  else haar_fwd_kernel<<<cdiv(M, 256), 256, 0, st>>>(in, out, N, C, h, w);
  SELFC_LAUNCH_CHECK("haar_kernel");
  return 0;
}

// ------------------------------------------------------------------------------------------------------
// Quantisation (Quantization.py:9-12): round-half-even of clamp(x,0,1)*255.
// ------------------------------------------------------------------------------------------------------
__device__ __forceinline__ float quant_code(float v) {
  v = fminf(fmaxf(v, 0.f), 1.f);
  return rintf(v * 255.0f);
}

__global__ void quantize_kernel(const float* __restrict__ x, uint8_t* __restrict__ q8, float* __restrict__ qf, size_t n) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float c = quant_code(x[i]);
  if (q8) q8[i] = (uint8_t)c;
  if (qf) qf[i] = c / 255.0f;
}

// ------------------------------------------------------------------------------------------------------
// Export of the forward result: planar z -> out51 [N,51,h,w] (+ LR quantised to u8 / fp32 grid).  One thread per
// pixel: both the planar reads and the NCHW writes are coalesced across the warp.
// ------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) export_down_kernel(const float* __restrict__ z, float* __restrict__ out51,
                                                          uint8_t* __restrict__ lr_u8, float* __restrict__ lr_q,
                                                          long long M, long long hw) {
  const long long m = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= M) return;
  const long long n = m / hw, pix = m % hw;
  const float4 x1 = load4(z + quad_off(M, 0, m));
  const float lr[3] = {x1.x, x1.y, x1.z};
  if (out51) {
    float* o = out51 + n * 51 * hw + pix;
#pragma unroll
    for (int c = 0; c < 3; ++c) o[c * hw] = lr[c];
#pragma unroll
    for (int q = 0; q < kSQuads; ++q) {
      const float4 r = load4(z + quad_off(M, 1 + q, m));
      o[(3 + 4 * q) * hw] = r.x; o[(4 + 4 * q) * hw] = r.y; o[(5 + 4 * q) * hw] = r.z; o[(6 + 4 * q) * hw] = r.w;
    }
  }
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    float code = quant_code(lr[c]);
    if (lr_u8) lr_u8[(n * 3 + c) * hw + pix] = (uint8_t)code;
    if (lr_q) lr_q[(n * 3 + c) * hw + pix] = code / 255.0f;
  }
}

// hf part of z -> [N,48,h,w]
__global__ void __launch_bounds__(256) export_hf_kernel(const float* __restrict__ z, float* __restrict__ hf, long long M, long long hw) {
  const long long m = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= M) return;
  const long long n = m / hw, pix = m % hw;
  float* o = hf + n * kHF * hw + pix;
#pragma unroll
  for (int q = 0; q < kSQuads; ++q) {
    const float4 r = load4(z + quad_off(M, 1 + q, m));
    o[(4 * q) * hw] = r.x; o[(4 * q + 1) * hw] = r.y; o[(4 * q + 2) * hw] = r.z; o[(4 * q + 3) * hw] = r.w;
  }
}

// ------------------------------------------------------------------------------------------------------
// NCHW [N,C,h,w] fp32 -> pixel-major T buffer [M][pitch] at channel offset `off`, zero-filling channels
// [C, cpad).  Used for the LR ingest of the reverse pass (x1 into z and the X slots of G/H/local_m1) and by
// the component entry points.
// ------------------------------------------------------------------------------------------------------
template <typename T>
__global__ void nchw_to_dense_kernel(const float* __restrict__ x, int ctot, int c0, T* __restrict__ dst, int pitch, long long slabM,
                                     int off, int C, int cpad, long long M, long long hw) {
  const long long m = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= M) return;
  const long long n = m / hw, pix = m % hw;
  for (int c = 0; c < cpad; ++c)
    store1(dst + dense_off(m, off + c, pitch, slabM), c < C ? __ldg(x + (n * ctot + c0 + c) * hw + pix) : 0.f);
}

// LR ingest of the reverse pass in ONE coalesced pass: lr [N,3,h,w] fp32 -> quad 0 of the latent state (x1 of the
// reversed block 8) and the 16-channel X slab (3 values + 13 zeros, two 16-byte stores) of up to three slab-planar dense
// buffers (G, H, local_m1).  The generic per-channel kernel above took 82 us per buffer at 1080p.
// X2 (BF16X3 mode, common.cuh): the slab row of a pixel is 64 bytes, [16 x hi | 16 x lo]
template <bool X2>
__global__ void __launch_bounds__(256) lr_ingest_slab_kernel(const float* __restrict__ lr, float* __restrict__ z, __nv_bfloat16* __restrict__ d0,
                                                             __nv_bfloat16* __restrict__ d1, __nv_bfloat16* __restrict__ d2, long long M,
                                                             long long hw) {
  const long long m = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= M) return;
  const long long n = m / hw, pix = m - n * hw;
  const float a = __ldg(lr + (n * 3 + 0) * hw + pix), b = __ldg(lr + (n * 3 + 1) * hw + pix), c = __ldg(lr + (n * 3 + 2) * hw + pix);
  store4(z + quad_off((size_t)M, 0, (size_t)m), make_float4(a, b, c, 0.f));
  __nv_bfloat162 ab = __floats2bfloat162_rn(a, b), c0 = __floats2bfloat162_rn(c, 0.f);
  uint4 lo = make_uint4(*reinterpret_cast<uint32_t*>(&ab), *reinterpret_cast<uint32_t*>(&c0), 0u, 0u);
  const uint4 zero = make_uint4(0u, 0u, 0u, 0u);
  __nv_bfloat16* dst[3] = {d0, d1, d2};
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    if (dst[i] == nullptr) continue;
    if (X2) {
      uint4* o = reinterpret_cast<uint4*>(dst[i] + m * 32);     // slab 0: [M][16 hi | 16 lo]
      o[0] = lo;
      o[1] = zero;
      o[2] = make_uint4(x2_pack_lo(a, b, lo.x), x2_pack_lo(c, 0.f, lo.y), 0u, 0u);
      o[3] = zero;
    } else {
      uint4* o = reinterpret_cast<uint4*>(dst[i] + m * 16);       // slab 0 of a slab-planar buffer: [M][16]
      o[0] = lo;
      o[1] = zero;
    }
  }
}

int launch_lr_ingest_slab(const float* lr, float* z, __nv_bfloat16* d0, __nv_bfloat16* d1, __nv_bfloat16* d2, long long M, long long hw,
                          cudaStream_t st, bool x2) {
  if (M == 0) return 0;
  if (x2) lr_ingest_slab_kernel<true><<<cdiv(M, 256), 256, 0, st>>>(lr, z, d0, d1, d2, M, hw);
  else lr_ingest_slab_kernel<false><<<cdiv(M, 256), 256, 0, st>>>(lr, z, d0, d1, d2, M, hw);
  SELFC_LAUNCH_CHECK("lr_ingest_slab_kernel");
  return 0;
}

template <typename T>
__global__ void dense_to_nchw_kernel(const T* __restrict__ src, int pitch, long long slabM, int off, float* __restrict__ y, int C,
                                     long long M, long long hw) {
  const long long m = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= M) return;
  const long long n = m / hw, pix = m % hw;
  for (int c = 0; c < C; ++c) y[(n * C + c) * hw + pix] = load1(src + dense_off(m, off + c, pitch, slabM));
}

// ------------------------------------------------------------------------------------------------------
// LR_ref of `distortion: sr_bd` (models/Guassian.py:7-52): reflect-pad 14, 13x13 Gaussian (sigma 1.6, the taps of
// scipy.ndimage.gaussian_filter on a dirac, passed in), stride 4, crop 2 px of the padded result on each side
// => out[i,j] = sum_{u,v<13} k[u,v] * x[reflect(4i+u-6), reflect(4j+v-6)].
// ------------------------------------------------------------------------------------------------------
__device__ __forceinline__ int reflect_idx(int p, int n) {
  if (p < 0) p = -p;
  if (p >= n) p = 2 * n - 2 - p;
  return p;
}
__global__ void __launch_bounds__(256) gaussian_down_kernel(const float* __restrict__ x, const float* __restrict__ k13,
                                                            float* __restrict__ y, int NC, int H, int W) {
  __shared__ float ks[169];
  for (int i = threadIdx.x; i < 169; i += blockDim.x) ks[i] = k13[i];
  __syncthreads();
  const int h = H / 4, w = W / 4;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)NC * h * w) return;
  const int j = (int)(idx % w), i = (int)((idx / w) % h);
  const long long nc = idx / ((long long)w * h);
  const float* src = x + nc * H * W;
  float acc = 0.f;
  for (int u = 0; u < 13; ++u) {
    const int yy = reflect_idx(4 * i + u - 6, H);
    for (int v = 0; v < 13; ++v) acc += ks[u * 13 + v] * __ldg(src + (long long)yy * W + reflect_idx(4 * j + v - 6, W));
  }
  y[idx] = acc;
}

int launch_gaussian_down(const float* x, const float* k13, float* y, int NC, int H, int W, cudaStream_t st) {
  const long long n = (long long)NC * (H / 4) * (W / 4);
  if (n == 0) return 0;
  gaussian_down_kernel<<<cdiv(n, 256), 256, 0, st>>>(x, k13, y, NC, H, W);
  SELFC_LAUNCH_CHECK("gaussian_down_kernel");
  return 0;
}

// ---- launchers ----------------------------------------------------------------------------------------
int launch_fa_fwd_nchw(const float* x, float* out51, int N, int h, int w, cudaStream_t st) {
  long long M = (long long)N * h * w;
  fa_fwd_kernel<true, float><<<cdiv(M, 256), 256, 0, st>>>(x, out51, nullptr, 0, 0, N, h, w);
  SELFC_LAUNCH_CHECK("fa_fwd_kernel<nchw>");
  return 0;
}

template <typename T>
int launch_fa_fwd_z(const float* x, float* z, T* fbuf, int fpitch, long long fslabM, int N, int h, int w, cudaStream_t st) {
  long long M = (long long)N * h * w;
  fa_fwd_kernel<false, T><<<cdiv(M, 256), 256, 0, st>>>(x, z, fbuf, fpitch, fslabM, N, h, w);
  SELFC_LAUNCH_CHECK("fa_fwd_kernel<z>");
  return 0;
}
template <typename T>
int launch_fa_fwd_z_u8(const uint8_t* x, float* z, T* fbuf, int fpitch, long long fslabM, int N, int h, int w, cudaStream_t st) {
  long long M = (long long)N * h * w;
  fa_fwd_kernel<false, T, true><<<cdiv(M, 256), 256, 0, st>>>(x, z, fbuf, fpitch, fslabM, N, h, w);
  SELFC_LAUNCH_CHECK("fa_fwd_kernel<z,u8>");
  return 0;
}
template int launch_fa_fwd_z_u8<float>(const uint8_t*, float*, float*, int, long long, int, int, int, cudaStream_t);
template int launch_fa_fwd_z_u8<__nv_bfloat16>(const uint8_t*, float*, __nv_bfloat16*, int, long long, int, int, int, cudaStream_t);
template int launch_fa_fwd_z_u8<bfx2>(const uint8_t*, float*, bfx2*, int, long long, int, int, int, cudaStream_t);

int launch_fa_rev_u8(const float* z, uint8_t* y, int N, int h, int w, cudaStream_t st) {
  long long M = (long long)N * h * w;
  fa_rev_kernel<false, true><<<cdiv(M, 256), 256, 0, st>>>(z, y, N, h, w);
  SELFC_LAUNCH_CHECK("fa_rev_kernel<u8>");
  return 0;
}

int launch_frames_from_u8(const uint8_t* img, float* x, long long N, long long HW, cudaStream_t st) {
  if (N * HW == 0) return 0;
  frames_from_u8_kernel<<<cdiv(N * HW, 256), 256, 0, st>>>(img, x, N, HW);
  SELFC_LAUNCH_CHECK("frames_from_u8_kernel");
  return 0;
}

int launch_frames_to_u8(const float* x, uint8_t* img, long long N, long long HW, cudaStream_t st) {
  if (N * HW == 0) return 0;
  frames_to_u8_kernel<<<cdiv(N * HW, 256), 256, 0, st>>>(x, img, N, HW);
  SELFC_LAUNCH_CHECK("frames_to_u8_kernel");
  return 0;
}
template int launch_fa_fwd_z<float>(const float*, float*, float*, int, long long, int, int, int, cudaStream_t);
template int launch_fa_fwd_z<__nv_bfloat16>(const float*, float*, __nv_bfloat16*, int, long long, int, int, int, cudaStream_t);
template int launch_fa_fwd_z<bfx2>(const float*, float*, bfx2*, int, long long, int, int, int, cudaStream_t);

int launch_fa_rev(const float* z, bool z_is_nchw, float* y, int N, int h, int w, cudaStream_t st) {
  long long M = (long long)N * h * w;
  if (z_is_nchw) fa_rev_kernel<true><<<cdiv(M, 256), 256, 0, st>>>(z, y, N, h, w);
  else fa_rev_kernel<false><<<cdiv(M, 256), 256, 0, st>>>(z, y, N, h, w);
  SELFC_LAUNCH_CHECK("fa_rev_kernel");
  return 0;
}

int launch_quantize(const float* x, uint8_t* q8, float* qf, size_t n, cudaStream_t st) {
  if (n == 0) return 0;
  quantize_kernel<<<cdiv((long long)n, 256), 256, 0, st>>>(x, q8, qf, n);
  SELFC_LAUNCH_CHECK("quantize_kernel");
  return 0;
}

int launch_export_down(const float* z, float* out51, uint8_t* lr_u8, float* lr_q, long long M, long long hw, cudaStream_t st) {
  export_down_kernel<<<cdiv(M, 256), 256, 0, st>>>(z, out51, lr_u8, lr_q, M, hw);
  SELFC_LAUNCH_CHECK("export_down_kernel");
  return 0;
}

int launch_export_hf(const float* z, float* hf, long long M, long long hw, cudaStream_t st) {
  export_hf_kernel<<<cdiv(M, 256), 256, 0, st>>>(z, hf, M, hw);
  SELFC_LAUNCH_CHECK("export_hf_kernel");
  return 0;
}

template <typename T>
int launch_nchw_to_dense(const float* x, T* dst, int pitch, long long slabM, int off, int C, int cpad, long long M, long long hw,
                         cudaStream_t st) {
  nchw_to_dense_kernel<T><<<cdiv(M, 256), 256, 0, st>>>(x, C, 0, dst, pitch, slabM, off, C, cpad, M, hw);
  SELFC_LAUNCH_CHECK("nchw_to_dense_kernel");
  return 0;
}
template <typename T>
int launch_nchw_slice_to_dense(const float* x, int ctot, int c0, T* dst, int pitch, long long slabM, int off, int C, int cpad,
                               long long M, long long hw, cudaStream_t st) {
  nchw_to_dense_kernel<T><<<cdiv(M, 256), 256, 0, st>>>(x, ctot, c0, dst, pitch, slabM, off, C, cpad, M, hw);
  SELFC_LAUNCH_CHECK("nchw_to_dense_kernel");
  return 0;
}
template int launch_nchw_slice_to_dense<float>(const float*, int, int, float*, int, long long, int, int, int, long long, long long,
                                               cudaStream_t);
template int launch_nchw_slice_to_dense<__nv_bfloat16>(const float*, int, int, __nv_bfloat16*, int, long long, int, int, int, long long,
                                                       long long, cudaStream_t);
template int launch_nchw_slice_to_dense<bfx2>(const float*, int, int, bfx2*, int, long long, int, int, int, long long, long long, cudaStream_t);
template int launch_nchw_to_dense<bfx2>(const float*, bfx2*, int, long long, int, int, int, long long, long long, cudaStream_t);
template int launch_nchw_to_dense<float>(const float*, float*, int, long long, int, int, int, long long, long long, cudaStream_t);
template int launch_nchw_to_dense<__nv_bfloat16>(const float*, __nv_bfloat16*, int, long long, int, int, int, long long, long long,
                                                 cudaStream_t);

template <typename T>
int launch_dense_to_nchw(const T* src, int pitch, long long slabM, int off, float* y, int C, long long M, long long hw, cudaStream_t st) {
  dense_to_nchw_kernel<T><<<cdiv(M, 256), 256, 0, st>>>(src, pitch, slabM, off, y, C, M, hw);
  SELFC_LAUNCH_CHECK("dense_to_nchw_kernel");
  return 0;
}
template int launch_dense_to_nchw<float>(const float*, int, long long, int, float*, int, long long, long long, cudaStream_t);
template int launch_dense_to_nchw<__nv_bfloat16>(const __nv_bfloat16*, int, long long, int, float*, int, long long, long long, cudaStream_t);
template int launch_dense_to_nchw<bfx2>(const bfx2*, int, long long, int, float*, int, long long, long long, cudaStream_t);

}  // namespace selfc
