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

// First statement of a kernel launched through launch_chain() (wgrad_tc.h): let the next kernel of the stream be scheduled, then wait
// until the previous kernel's results are visible.  Without the launch attribute both instructions do nothing.
__device__ __forceinline__ void chain_entry() {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");
}

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

// ---- BF16X3 mode: activations as (hi, lo) bf16 pairs ------------------------------------------------------------
// The fp32-gate mode on tensor cores (BASELINE configs[1]) keeps every activation as hi = bf16(v), lo = bf16(v - hi) (16 mantissa
// bits) and every weight the same way, and forms A.W = A_hi.W_hi + A_hi.W_lo + A_lo.W_hi with three bf16 MMAs into one fp32
// accumulator (the dropped A_lo.W_lo term is 2^-18 relative).  Storage: the buffers keep the layouts of common.cuh's dense
// buffers with a 4-byte element, but the two halves of an element are NOT adjacent: every run of 16 consecutive elements (one
// 64-byte row: 16 channels of one pixel, in the slab-planar and in the pixel-major layout alike) is stored as [16 x hi | 16 x lo],
// i.e. as two K = 16 operand rows of 32 bytes that the MMA descriptors address directly (+0 / +32 bytes inside a SWIZZLE_64B or
// SWIZZLE_128B row).  `bfx2*` is therefore a HANDLE: pointer arithmetic counts logical elements, and the helpers below turn
// the handle into the addresses of the two halves (buffers are 64-byte aligned; rows never straddle).
struct bfx2 { uint32_t opaque; };
__host__ __device__ __forceinline__ size_t x2_hi_index(size_t logical) { return ((logical >> 4) << 5) + (logical & 15); }   // in bf16 units; lo = +16
__device__ __forceinline__ __nv_bfloat16* x2_hi_ptr(const bfx2* p) {
  const uintptr_t a = reinterpret_cast<uintptr_t>(p);
  return reinterpret_cast<__nv_bfloat16*>((a & ~(uintptr_t)63) + ((a & 63) >> 1));
}
__device__ __forceinline__ void x2_split(float v, __nv_bfloat16& hi, __nv_bfloat16& lo) {
  hi = __float2bfloat16_rn(v);
  lo = __float2bfloat16_rn(v - __bfloat162float(hi));
}
__device__ __forceinline__ uint32_t x2_pack_hi(float a, float b) {
  __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ uint32_t x2_pack_lo(float a, float b, uint32_t hi) {      // residuals of the two values packed in `hi`
  return x2_pack_hi(a - __uint_as_float(hi << 16), b - __uint_as_float(hi & 0xffff0000u));
}
// 8 consecutive channels (multiple of 8 inside a 16-row) -> 16 bytes of hi + 16 bytes of lo at +32 bytes
__device__ __forceinline__ void x2_store8(__nv_bfloat16* hi_ptr, const float* v) {
  uint4 h, l;
  h.x = x2_pack_hi(v[0], v[1]); h.y = x2_pack_hi(v[2], v[3]); h.z = x2_pack_hi(v[4], v[5]); h.w = x2_pack_hi(v[6], v[7]);
  l.x = x2_pack_lo(v[0], v[1], h.x); l.y = x2_pack_lo(v[2], v[3], h.y); l.z = x2_pack_lo(v[4], v[5], h.z); l.w = x2_pack_lo(v[6], v[7], h.w);
  *reinterpret_cast<uint4*>(hi_ptr) = h;
  *reinterpret_cast<uint4*>(hi_ptr + 16) = l;
}
__device__ __forceinline__ void x2_load8(const __nv_bfloat16* hi_ptr, float* v) {
  const uint4 h = *reinterpret_cast<const uint4*>(hi_ptr);
  const uint4 l = *reinterpret_cast<const uint4*>(hi_ptr + 16);
  const uint32_t hw[4] = {h.x, h.y, h.z, h.w}, lw[4] = {l.x, l.y, l.z, l.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    v[2 * i] = __uint_as_float(hw[i] << 16) + __uint_as_float(lw[i] << 16);
    v[2 * i + 1] = __uint_as_float(hw[i] & 0xffff0000u) + __uint_as_float(lw[i] & 0xffff0000u);
  }
}
__device__ __forceinline__ float4 load4(const bfx2* p) {
  const __nv_bfloat16* h = x2_hi_ptr(p);
  const uint2 a = *reinterpret_cast<const uint2*>(h);
  const uint2 b = *reinterpret_cast<const uint2*>(h + 16);
  return make_float4(__uint_as_float(a.x << 16) + __uint_as_float(b.x << 16),
                     __uint_as_float(a.x & 0xffff0000u) + __uint_as_float(b.x & 0xffff0000u),
                     __uint_as_float(a.y << 16) + __uint_as_float(b.y << 16),
                     __uint_as_float(a.y & 0xffff0000u) + __uint_as_float(b.y & 0xffff0000u));
}
__device__ __forceinline__ void store4(bfx2* p, float4 v) {
  __nv_bfloat16* h = x2_hi_ptr(p);
  uint2 a, b;
  a.x = x2_pack_hi(v.x, v.y); a.y = x2_pack_hi(v.z, v.w);
  b.x = x2_pack_lo(v.x, v.y, a.x); b.y = x2_pack_lo(v.z, v.w, a.y);
  *reinterpret_cast<uint2*>(h) = a;
  *reinterpret_cast<uint2*>(h + 16) = b;
}
// single elements (the layout kernels of the component entry points)
__device__ __forceinline__ float load1(const float* p) { return *p; }
__device__ __forceinline__ float load1(const __nv_bfloat16* p) { return __bfloat162float(*p); }
__device__ __forceinline__ float load1(const bfx2* p) {
  const __nv_bfloat16* h = x2_hi_ptr(p);
  return __bfloat162float(h[0]) + __bfloat162float(h[16]);
}
__device__ __forceinline__ void store1(float* p, float v) { *p = v; }
__device__ __forceinline__ void store1(__nv_bfloat16* p, float v) { *p = __float2bfloat16_rn(v); }
__device__ __forceinline__ void store1(bfx2* p, float v) {
  __nv_bfloat16* h = x2_hi_ptr(p);
  x2_split(v, h[0], h[16]);
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
