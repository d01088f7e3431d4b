// Shared declarations for the selfc_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>

#include "../../include/selfc_b200.h"

namespace selfc {

constexpr int kHF = 48;         // 3 * 4 * 4 high-frequency channels
constexpr int kGmmK = 5;
constexpr int kGrowth = 32;     // D2DTInput gc
// Latent state z: PLANAR quads [13][M][4] fp32 -- quad 0 = (x1_0, x1_1, x1_2, 0) the LR part, quad 1+q = x2[4q..4q+3]
// the HF part.  One thread per pixel then reads / writes 16 bytes per quad with consecutive lanes on consecutive
// addresses (coalesced), which is what every epilogue and layout kernel does.  The coupling log-scale s is [12][M][4].
constexpr int kZQuads = 13;
constexpr int kSQuads = 12;
__host__ __device__ __forceinline__ size_t quad_off(size_t M, int quad, size_t m) { return ((size_t)quad * M + m) * 4; }
constexpr int kStpC = 64;

// Dense-block buffers ("dense buffers") hold the concatenation [X | x1 | x2 | x3 | x4] of a D2DTInput.
//   pixel-major (fp32 mode):  element (m, c) at  m * pitch + c
//   slab-planar (bf16 mode):  16-channel slabs [pitch/16][M][16]: element (m, c) at ((c >> 4) * M + m) * 16 + (c & 15)
// In the slab layout every access of every kernel is contiguous along the pixel index: a conv tile row of 32 pixels of
// one K-step is one 1 KB run, a conv reads exactly the slabs it consumes (no over-fetch of the channels it does not),
// and an epilogue's 32 output channels are two 32-byte stores per pixel with consecutive lanes on consecutive addresses.
// `slabM` = M selects the slab layout, 0 the pixel-major one.  Groups of 4/8/16 channels starting at a multiple of
// their size never straddle a slab.
__host__ __device__ __forceinline__ size_t dense_off(long long m, int c, int pitch, long long slabM) {
  return slabM ? ((size_t)(c >> 4) * (size_t)slabM + (size_t)m) * 16 + (size_t)(c & 15) : (size_t)m * pitch + c;
}

// ---- error plumbing ---------------------------------------------------------------------------------
void set_error(const char* fmt, ...);
uint64_t& launch_counter();

#define SELFC_CHECK_ARG(cond, ...)            \
  do {                                        \
    if (!(cond)) {                            \
      selfc::set_error(__VA_ARGS__);          \
      return SELFC_E_ARG;                     \
    }                                         \
  } while (0)

#define SELFC_CUDA(expr)                                                                   \
  do {                                                                                     \
    cudaError_t e__ = (expr);                                                              \
    if (e__ != cudaSuccess) {                                                              \
      selfc::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), __FILE__, __LINE__); \
      return SELFC_E_CUDA;                                                                 \
    }                                                                                      \
  } while (0)

#define SELFC_LAUNCH_CHECK(name)                                                           \
  do {                                                                                     \
    selfc::launch_counter()++;                                                             \
    cudaError_t e__ = cudaGetLastError();                                                  \
    if (e__ != cudaSuccess) {                                                              \
      selfc::set_error("launch of %s failed: %s", name, cudaGetErrorString(e__));          \
      return SELFC_E_CUDA;                                                                 \
    }                                                                                      \
  } while (0)

#define SELFC_TRY(expr)        \
  do {                         \
    int rc__ = (expr);         \
    if (rc__ != 0) return rc__; \
  } while (0)

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }
inline int cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }
inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// ---- element helpers --------------------------------------------------------------------------------
__device__ __forceinline__ float to_f(float v) { return v; }
__device__ __forceinline__ float to_f(__nv_bfloat16 v) { return __bfloat162float(v); }
template <typename T> __device__ __forceinline__ T from_f(float v);
template <> __device__ __forceinline__ float from_f<float>(float v) { return v; }
template <> __device__ __forceinline__ __nv_bfloat16 from_f<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

__device__ __forceinline__ float lrelu02(float v) { return v > 0.f ? v : 0.2f * v; }

// 4 consecutive elements -> float4
__device__ __forceinline__ float4 load4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ float4 load4(const __nv_bfloat16* p) {
  uint2 r = *reinterpret_cast<const uint2*>(p);
  __nv_bfloat162 a = *reinterpret_cast<__nv_bfloat162*>(&r.x);
  __nv_bfloat162 b = *reinterpret_cast<__nv_bfloat162*>(&r.y);
  return make_float4(__low2float(a), __high2float(a), __low2float(b), __high2float(b));
}
__device__ __forceinline__ void store4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
__device__ __forceinline__ void store4(__nv_bfloat16* p, float4 v) {
  __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y);
  __nv_bfloat162 b = __floats2bfloat162_rn(v.z, v.w);
  uint2 r;
  r.x = *reinterpret_cast<uint32_t*>(&a);
  r.y = *reinterpret_cast<uint32_t*>(&b);
  *reinterpret_cast<uint2*>(p) = r;
}

// ---- Philox4x32-10 counter RNG (keyed on the reference's eps linear index; SURVEY 7.2 "RNG") ---------
__host__ __device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                                       uint32_t k0, uint32_t k1, uint32_t out[4]) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    uint64_t p0 = (uint64_t)M0 * c0;
    uint64_t p1 = (uint64_t)M1 * c2;
    uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
    uint32_t n1 = (uint32_t)p1;
    uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
    uint32_t n3 = (uint32_t)p0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += W0; k1 += W1;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

// Noise stream (DESIGN.md "Noise"): the reference's eps tensor is [B,48,5,T,h,w] (SelfC_GMM_arch_inv.py:412-415).  One
// Philox4x32-10 call, counter = the linear index g of (b, hf/4, k, t, pixel) in [B,12,5,T,h*w], yields the FOUR normals
// of hf = 4*(hf/4) + 0..3: Box-Muller on words (0,1) -> cos, sin branches, on words (2,3) -> cos, sin branches.
// So a thread-per-pixel sampler pays one Philox per (k, hf-quad), and the values depend only on (seed, offset) and
// the element's coordinates, never on tiling, launch shape or GPU count.
__device__ __forceinline__ void philox_normal4(uint64_t g, uint64_t seed, uint64_t offset, float out[4]) {
  uint32_t r[4];
  philox4x32_10((uint32_t)g, (uint32_t)(g >> 32), (uint32_t)offset, (uint32_t)(offset >> 32),
                (uint32_t)seed, (uint32_t)(seed >> 32), r);
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const float u1 = ((float)r[2 * h] + 1.0f) * 2.3283064365386963e-10f;   // (0, 1]
    const float u2 = (float)r[2 * h + 1] * 2.3283064365386963e-10f;        // [0, 1]
#ifdef SELFC_FAST_NORMAL
    // hardware approximations (MUFU.LG2 / SIN / COS, ~2^-21 relative): 40 % fewer instructions in the issue-bound sampler
    const float rad = __fsqrt_rn(-2.0f * __logf(u1));
    float sn, cs;
    __sincosf(6.283185307179586f * u2, &sn, &cs);
#else
    const float rad = sqrtf(-2.0f * logf(u1));
    float sn, cs;
    sincospif(2.0f * u2, &sn, &cs);
#endif
    out[2 * h] = rad * cs;
    out[2 * h + 1] = rad * sn;
  }
}
__host__ __device__ __forceinline__ uint64_t eps_group(long long b, int hfq, int k, int t, long long pix, int T, long long hw) {
  return (uint64_t)(((((b * (kHF / 4) + hfq) * kGmmK + k) * T + t) * hw) + pix);
}
// single element eps[b, hf, k, t, pix] (the non-hot callers)
__device__ __forceinline__ float philox_eps(long long b, int hf, int k, int t, long long pix, int T, long long hw, uint64_t seed,
                                            uint64_t offset) {
  float n4[4];
  philox_normal4(eps_group(b, hf >> 2, k, t, pix, T, hw), seed, offset, n4);
  return n4[hf & 3];
}

}  // namespace selfc
