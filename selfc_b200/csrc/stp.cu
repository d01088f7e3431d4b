// STPNet pieces that are not convolutions: GlobalAgg statistics (pooled descriptor -> T x T mixing matrix) and
// the soft-GMM sampler with counter-based noise.
//
// Reference behaviour restated (not copied): models/modules/SelfC_GMM_arch_inv.py:257-285 (GlobalAgg),
// :383-394 + :412-417 (sampler / reparametrize).
#include <cstdlib>
#include <type_traits>
#include <cuda_fp16.h>
#include <cuda_pipeline.h>
#include "common.cuh"
#include "kernels.h"

namespace selfc {

// ------------------------------------------------------------------------------------------------------
// fc(adaptive_avg_pool2d(x,(32,32))) is linear in x: d[n,ch] = fcb + sum_pix wmap[pix] * x[n,pix,ch] with
// wmap[y,x] = sum over the pooling bins (i,j) that contain (y,x) of fcw[i*32+j] / (|bin_i| * |bin_j|).
// Bins follow PyTorch: rows floor(i*h/32) .. ceil((i+1)*h/32)-1 (they overlap when h/32 is not integral).
// ------------------------------------------------------------------------------------------------------
__global__ void ga_wmap_kernel(const float* __restrict__ fcw, float* __restrict__ wmap, int h, int w) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= h * w) return;
  const int y = p / w, x = p % w;
  float acc = 0.f;
  for (int i = 0; i < 32; ++i) {
    const int ys = (i * h) / 32, ye = ((i + 1) * h + 31) / 32;
    if (y < ys || y >= ye) continue;
    for (int j = 0; j < 32; ++j) {
      const int xs = (j * w) / 32, xe = ((j + 1) * w + 31) / 32;
      if (x < xs || x >= xe) continue;
      acc += __ldg(fcw + i * 32 + j) / (float)((ye - ys) * (xe - xs));
    }
  }
  wmap[p] = acc;
}

// partial[n][split][64] = sum over this split's pixels of wmap[p] * x[n,p,ch]   (deterministic two-stage sum)
template <typename T>
__global__ void __launch_bounds__(256) ga_stat_kernel(const T* __restrict__ x, int pitch, const float* __restrict__ wmap,
                                                      float* __restrict__ partial, int nsplit, int hw) {
  const int split = blockIdx.x, n = blockIdx.y;
  const int per = (hw + nsplit - 1) / nsplit;
  const int p0 = split * per, p1 = min(hw, p0 + per);
  if constexpr (std::is_same<T, __nv_bfloat16>::value) {
    // bf16 features: 8 threads x 16 bytes per pixel, 32 pixels per pass, four passes (64 bytes per thread) in flight
    __shared__ float red[32][64 + 4];
    const int ch = (threadIdx.x & 7) * 8;
    const int lane = threadIdx.x >> 3;
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    const T* xb = x + (long long)n * hw * pitch + ch;
    for (int p = p0 + lane; p < p1; p += 128) {
      uint4 r[4];
      float wv[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int pp = p + 32 * u;
        const bool ok = pp < p1;
        wv[u] = ok ? __ldg(wmap + pp) : 0.f;
        r[u] = ok ? __ldg(reinterpret_cast<const uint4*>(xb + (long long)pp * pitch)) : make_uint4(0u, 0u, 0u, 0u);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const uint32_t w32[4] = {r[u].x, r[u].y, r[u].z, r[u].w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          acc[2 * j] = fmaf(wv[u], __uint_as_float(w32[j] << 16), acc[2 * j]);             // bf16 -> fp32 is a 16-bit shift
          acc[2 * j + 1] = fmaf(wv[u], __uint_as_float(w32[j] & 0xffff0000u), acc[2 * j + 1]);
        }
      }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) red[lane][ch + j] = acc[j];
    __syncthreads();
    if (threadIdx.x < 64) {
      float s = 0.f;
#pragma unroll
      for (int l = 0; l < 32; ++l) s += red[l][threadIdx.x];
      partial[((long long)n * nsplit + split) * 64 + threadIdx.x] = s;
    }
    return;
  } else {
  __shared__ float red[16][64 + 4];
  const int ch = (threadIdx.x & 15) * 4;
  const int lane = threadIdx.x >> 4;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int p = p0 + lane; p < p1; p += 16) {
    const float wv = __ldg(wmap + p);
    const float4 r = load4(x + ((long long)n * hw + p) * pitch + ch);
    acc.x += wv * r.x; acc.y += wv * r.y; acc.z += wv * r.z; acc.w += wv * r.w;
  }
  red[lane][ch] = acc.x; red[lane][ch + 1] = acc.y; red[lane][ch + 2] = acc.z; red[lane][ch + 3] = acc.w;
  __syncthreads();
  if (threadIdx.x < 64) {
    float s = 0.f;
#pragma unroll
    for (int l = 0; l < 16; ++l) s += red[l][threadIdx.x];
    partial[((long long)n * nsplit + split) * 64 + threadIdx.x] = s;
  }
  }
}

// one CTA of 1024 threads per clip: d (sum of the ga_stat partials) -> q,k -> A = q k^T / 64 -> row softmax -> W [T][T], column sums.
// A latency chain on one SM, so every stage is laid out for few, wide steps: proj2/proj3 weights are staged into shared memory
// with coalesced loads (row-per-thread __ldg reads cost 32 L1 wavefronts per instruction: 30 of the 43 us this kernel used to
// take) while the partial sums are in flight; each frame's partials are summed by S = 16/T thread groups (fixed order).
constexpr int kGaWPitch = 65;         // padded rows: thread c walks row c, the 32 lanes of a warp hit 32 different banks
__global__ void __launch_bounds__(1024) ga_weights_kernel(const float* __restrict__ partial, int nsplit, const float* __restrict__ fcb,
                                                          const float* __restrict__ p2w, const float* __restrict__ p2b,
                                                          const float* __restrict__ p3w, const float* __restrict__ p3b,
                                                          float* __restrict__ wmat, float* __restrict__ wsum, int T) {
  constexpr int MAXT = 32;
  extern __shared__ float sm[];
  float* d = sm;                      // [T][64]
  float* q = d + T * 64;              // [T][65]  (padded: the q.k products read rows with stride 65)
  float* k = q + T * 65;              // [T][65]
  float* A = k + T * 65;              // [T][MAXT]
  float* dpart = A + T * MAXT;        // [16][64]   per-group partial sums of stage 1
  float* w2s = dpart + 16 * 64;       // [64][65]
  float* w3s = w2s + 64 * kGaWPitch;  // [64][65]
  const int b = blockIdx.x;
  const int c = threadIdx.x & 63, grp = threadIdx.x >> 6;     // 16 groups of 64 threads
  // stage 0: weights -> shared memory (4096 + 4096 coalesced 4-byte cp.async, in flight while stage 1 waits for its partials)
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    const int e = u * 1024 + threadIdx.x;       // element (row e / 64, column e % 64)
    __pipeline_memcpy_async(w2s + (e >> 6) * kGaWPitch + (e & 63), p2w + e, 4);
    __pipeline_memcpy_async(w3s + (e >> 6) * kGaWPitch + (e & 63), p3w + e, 4);
  }
  __pipeline_commit();
  // stage 1: d[t][c] = fcb + sum over the nsplit partials, S groups per frame
  const int S = T <= 8 ? 16 / T : 1;
  const int per = (nsplit + S - 1) / S;
  for (int t0 = 0; t0 < T; t0 += 16 / S) {
    const int t = t0 + grp / S, sub = grp % S;
    float s = 0.f;
    if (t < T && grp < (16 / S) * S) {
      const int sp1 = min(nsplit, (sub + 1) * per);
      int sp = sub * per;
      const float* pp = partial + (((long long)b * T + t) * nsplit) * 64 + c;
      for (; sp + 16 <= sp1; sp += 16) {        // 16 independent loads in flight, summed in a fixed order (deterministic)
        float v[16];
#pragma unroll
        for (int u = 0; u < 16; ++u) v[u] = __ldg(pp + (long long)(sp + u) * 64);
#pragma unroll
        for (int u = 0; u < 16; ++u) s += v[u];
      }
      for (; sp < sp1; ++sp) s += __ldg(pp + (long long)sp * 64);
    }
    dpart[grp * 64 + c] = s;
    __syncthreads();
    if (sub == 0 && t < T && grp < (16 / S) * S) {
      float tot = dpart[grp * 64 + c];
      for (int u = 1; u < S; ++u) tot += dpart[(grp + u) * 64 + c];
      d[t * 64 + c] = tot + fcb[0];
    }
    __syncthreads();
  }
  __pipeline_wait_prior(0);
  __syncthreads();
  // stage 2: q = proj2(d), k = proj3(d): one thread per (q|k, frame, channel)
  for (int item = threadIdx.x; item < 2 * T * 64; item += blockDim.x) {
    const int which = item / (T * 64), t = (item / 64) % T;
    const float* wrow = (which ? w3s : w2s) + c * kGaWPitch;
    float acc = which ? p3b[c] : p2b[c];
#pragma unroll 16
    for (int i = 0; i < 64; ++i) acc += wrow[i] * d[t * 64 + i];
    (which ? k : q)[t * 65 + c] = acc;
  }
  __syncthreads();
  for (int e = threadIdx.x; e < T * T; e += blockDim.x) {
    const int t = e / T, u = e % T;
    float s = 0.f;
    for (int i = 0; i < 64; ++i) s += q[t * 65 + i] * k[u * 65 + i];
    A[t * MAXT + u] = s / 64.0f;
  }
  __syncthreads();
  if (threadIdx.x < T) {
    const int t = threadIdx.x;
    float mx = -INFINITY;
    for (int u = 0; u < T; ++u) mx = fmaxf(mx, A[t * MAXT + u]);
    float sum = 0.f;
    for (int u = 0; u < T; ++u) { A[t * MAXT + u] = expf(A[t * MAXT + u] - mx); sum += A[t * MAXT + u]; }
    for (int u = 0; u < T; ++u) {
      A[t * MAXT + u] = A[t * MAXT + u] / sum;
      wmat[((long long)b * T + t) * T + u] = A[t * MAXT + u];
    }
  }
  __syncthreads();
  if (threadIdx.x < T) {
    float s = 0.f;
    for (int t = 0; t < T; ++t) s += A[t * MAXT + threadIdx.x];
    wsum[(long long)b * T + threadIdx.x] = s;
  }
}

// ------------------------------------------------------------------------------------------------------
// GlobalAgg apply (BF16 mode): out[b,t',.] = x[b,t',.] + sum_t W[b,t,t'] * P[b,t,.]  with P = proj1(x) + bias computed by
// the tcgen05 pointwise GEMM (SelfC_GMM_arch_inv.py:266,278-285).  49 FMAs per output element are nothing for the GPU
// as a whole but were 205 us when done by the four epilogue warps per SM of the GEMM kernel; here they run at full
// occupancy: one thread per (pixel, 8-channel group), 16-byte loads / stores, consecutive lanes on consecutive addresses.
// ------------------------------------------------------------------------------------------------------
// 8 consecutive channels of a bf16 buffer at logical element offset `off` (a multiple of 8); X2: the buffer holds (hi, lo) pairs
// as 64-byte rows [16 x hi | 16 x lo] (BF16X3 mode, common.cuh) and the value is hi + lo
template <bool X2>
__device__ __forceinline__ void mix_load8(const __nv_bfloat16* base, size_t off, float* v) {
  if constexpr (X2) {
    x2_load8(base + x2_hi_index(off), v);
  } else {
    const uint4 r = __ldg(reinterpret_cast<const uint4*>(base + off));
    const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      v[2 * i] = __uint_as_float(w[i] << 16);
      v[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
    }
  }
}
template <bool X2>
__device__ __forceinline__ void mix_store8(__nv_bfloat16* base, size_t off, const float* v) {
  if constexpr (X2) {
    x2_store8(base + x2_hi_index(off), v);
  } else {
    uint4 r;
    __nv_bfloat162 h0 = __floats2bfloat162_rn(v[0], v[1]), h1 = __floats2bfloat162_rn(v[2], v[3]);
    __nv_bfloat162 h2 = __floats2bfloat162_rn(v[4], v[5]), h3 = __floats2bfloat162_rn(v[6], v[7]);
    r.x = *reinterpret_cast<uint32_t*>(&h0); r.y = *reinterpret_cast<uint32_t*>(&h1);
    r.z = *reinterpret_cast<uint32_t*>(&h2); r.w = *reinterpret_cast<uint32_t*>(&h3);
    *reinterpret_cast<uint4*>(base + off) = r;
  }
}

template <bool X2>
__global__ void __launch_bounds__(256) ga_mix_kernel(const __nv_bfloat16* __restrict__ P, const __nv_bfloat16* __restrict__ x,
                                                     const float* __restrict__ wmat, __nv_bfloat16* __restrict__ outT, int outT_pitch,
                                                     long long outT_slabM, float* __restrict__ outF, int outF_pitch,
                                                     __nv_bfloat16* __restrict__ outAct, int T, long long hw, long long M) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long m = idx >> 3;
  const int c0 = (int)(idx & 7) * 8;
  if (m >= M) return;
  const long long n = m / hw, pix = m - n * hw;
  const int tq = (int)(n % T);
  const long long b = n / T;
  float acc[8];
  mix_load8<X2>(x, (size_t)m * kStpC + c0, acc);
  const float* wm = wmat + b * T * T + tq;          // W[b][t][t' = tq]
  for (int t = 0; t < T; ++t) {
    const float wv = __ldg(wm + t * T);
    float pv[8];
    mix_load8<X2>(P, (size_t)((b * T + t) * hw + pix) * kStpC + c0, pv);
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = fmaf(wv, pv[j], acc[j]);
  }
  if (outT) mix_store8<X2>(outT, dense_off(m, c0, outT_pitch, outT_slabM), acc);
  if (outF) {
    float* o = outF + m * outF_pitch + c0;
    store4(o, make_float4(acc[0], acc[1], acc[2], acc[3]));
    store4(o + 4, make_float4(acc[4], acc[5], acc[6], acc[7]));
  }
  if (outAct) {
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = lrelu02(acc[j]);
    mix_store8<X2>(outAct, (size_t)m * kStpC + c0, acc);
  }
}

// T <= 8: one thread per (clip, pixel, 8-channel group) produces ALL T output frames, so every P element is read from
// DRAM exactly once (the per-output-frame kernel above re-reads each P value T times from frames that are 16 MB apart).
// 8 channels = 16-byte accesses: a warp reads 4 pixels x 128 bytes contiguously and writes 128-byte runs into each 16-channel
// slab of the dense buffer (the 4-channel form wrote 64-byte runs: 3.7 TB/s); 2 CTAs of 256 threads per SM keep 2 x T x 16 bytes
// per thread in flight.
template <bool X2>
__global__ void __launch_bounds__(256, 2) ga_mix_allframes_kernel(const __nv_bfloat16* __restrict__ P, const __nv_bfloat16* __restrict__ x,
                                                                  const float* __restrict__ wmat, __nv_bfloat16* __restrict__ outT,
                                                                  int outT_pitch, long long outT_slabM, float* __restrict__ outF,
                                                                  int outF_pitch, __nv_bfloat16* __restrict__ outAct, int T, long long hw,
                                                                  long long Bhw) {
  constexpr int TM = 8;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long bp = idx >> 3;                  // b * hw + pix
  const int c0 = (int)(idx & 7) * 8;
  if (bp >= Bhw) return;
  const long long b = bp / hw, pix = bp - b * hw;
  // raw 16-byte loads are kept packed until they are used (64 registers for P and x of all frames); X2 keeps P as fp32 sums of the
  // two halves and reads x when its output frame is produced
  uint4 pr[TM], xr[TM];
  float pf[X2 ? TM : 1][8];
#pragma unroll
  for (int t = 0; t < TM; ++t) {
    if (t < T) {
      const long long m = (b * T + t) * hw + pix;
      if constexpr (X2) {
        mix_load8<true>(P, (size_t)m * kStpC + c0, pf[t]);
      } else {
        pr[t] = __ldg(reinterpret_cast<const uint4*>(P + m * kStpC + c0));
        xr[t] = __ldg(reinterpret_cast<const uint4*>(x + m * kStpC + c0));
      }
    }
  }
  auto unpack = [](const uint4 r, float* v) {
    const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      v[2 * i] = __uint_as_float(w[i] << 16);
      v[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
    }
  };
  const float* wm = wmat + b * T * T;               // W[b][t][t']
#pragma unroll
  for (int tq = 0; tq < TM; ++tq) {
    if (tq >= T) continue;
    const long long m = (b * T + tq) * hw + pix;
    float acc[8];
    if constexpr (X2) mix_load8<true>(x, (size_t)m * kStpC + c0, acc);
    else unpack(xr[tq], acc);
#pragma unroll
    for (int t = 0; t < TM; ++t) {
      if (t < T) {
        const float wv = __ldg(wm + t * T + tq);
        float pv[8];
        if constexpr (X2) {
#pragma unroll
          for (int j = 0; j < 8; ++j) pv[j] = pf[t][j];
        } else {
          unpack(pr[t], pv);
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] = fmaf(wv, pv[j], acc[j]);
      }
    }
    if (outT) mix_store8<X2>(outT, dense_off(m, c0, outT_pitch, outT_slabM), acc);
    if (outF) {
      store4(outF + m * outF_pitch + c0, make_float4(acc[0], acc[1], acc[2], acc[3]));
      store4(outF + m * outF_pitch + c0 + 4, make_float4(acc[4], acc[5], acc[6], acc[7]));
    }
    if (outAct) {
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] = lrelu02(acc[j]);
      mix_store8<X2>(outAct, (size_t)m * kStpC + c0, acc);
    }
  }
}

int launch_ga_mix(const __nv_bfloat16* P, const __nv_bfloat16* x, const float* wmat, __nv_bfloat16* outT, int outT_pitch,
                  long long outT_slabM, float* outF, int outF_pitch, __nv_bfloat16* outAct, int B, int T, long long hw, cudaStream_t st,
                  bool x2) {
  const long long M = (long long)B * T * hw;
  if (M == 0) return 0;
  SELFC_CHECK_ARG(outF == nullptr || outF_pitch % 4 == 0, "ga_mix: outF pitch");
  if (T <= 8) {
    if (x2)
      ga_mix_allframes_kernel<true><<<cdiv((long long)B * hw * 8, 256), 256, 0, st>>>(P, x, wmat, outT, outT_pitch, outT_slabM, outF,
                                                                                     outF_pitch, outAct, T, hw, (long long)B * hw);
    else
      ga_mix_allframes_kernel<false><<<cdiv((long long)B * hw * 8, 256), 256, 0, st>>>(P, x, wmat, outT, outT_pitch, outT_slabM, outF,
                                                                                      outF_pitch, outAct, T, hw, (long long)B * hw);
    SELFC_LAUNCH_CHECK("ga_mix_allframes_kernel");
    return 0;
  }
  if (x2) ga_mix_kernel<true><<<cdiv(M * 8, 256), 256, 0, st>>>(P, x, wmat, outT, outT_pitch, outT_slabM, outF, outF_pitch, outAct, T, hw, M);
  else ga_mix_kernel<false><<<cdiv(M * 8, 256), 256, 0, st>>>(P, x, wmat, outT, outT_pitch, outT_slabM, outF, outF_pitch, outAct, T, hw, M);
  SELFC_LAUNCH_CHECK("ga_mix_kernel");
  return 0;
}

// ------------------------------------------------------------------------------------------------------
// Soft-GMM sampler: one warp per LR pixel.  Channel ch = hf*15 + k*3 + j (j: 0 logit, 1 log-scale, 2 mean);
// the softmax runs over the 48 hf channels for each k; v[hf] = sum_k pi * (eps*exp(clamp(ls,-7,7)) + mu).
// eps is read from `eps` (reference layout [B,48,5,T,h,w]) or generated by Philox at that linear index.
// ------------------------------------------------------------------------------------------------------
template <bool P_NCHW, bool V_NCHW>
__global__ void __launch_bounds__(128) gmm_sample_kernel(const float* __restrict__ params, const float* __restrict__ eps, uint64_t seed,
                                                         uint64_t offset, float* __restrict__ v, int vpitch, int voff, int T, int h,
                                                         int w, long long M) {
  __shared__ __align__(16) float sp[4][720];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long m = (long long)blockIdx.x * 4 + warp;
  if (m >= M) return;
  const long long hw = (long long)h * w;
  const long long n = m / hw, pix = m - n * hw;
  const int t = (int)(n % T);
  const long long b = n / T;
  float* P = sp[warp];
  if (P_NCHW) {
    for (int c = lane; c < 720; c += 32) P[c] = __ldg(params + (n * 720 + c) * hw + pix);
  } else {
    const float4* src = reinterpret_cast<const float4*>(params + m * 720);
    for (int c = lane; c < 180; c += 32) reinterpret_cast<float4*>(P)[c] = __ldg(src + c);
  }
  __syncwarp();
  // each lane owns hf = lane and (lane < 16) hf = lane + 32
  const int hf0 = lane, hf1 = lane + 32;
  const bool has1 = hf1 < kHF;
  float out0 = 0.f, out1 = 0.f;
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
    {
      const float ls = fminf(fmaxf(P[hf0 * 15 + k * 3 + 1], -7.f), 7.f);
      const float mu = P[hf0 * 15 + k * 3 + 2];
      const uint64_t idx = (uint64_t)(((((b * kHF + hf0) * kGmmK + k) * T + t) * hw) + pix);
      const float ep = eps ? __ldg(eps + idx) : philox_eps(b, hf0, k, t, pix, T, hw, seed, offset);
      out0 += (e0 / sum) * (ep * expf(ls) + mu);
    }
    if (has1) {
      const float ls = fminf(fmaxf(P[hf1 * 15 + k * 3 + 1], -7.f), 7.f);
      const float mu = P[hf1 * 15 + k * 3 + 2];
      const uint64_t idx = (uint64_t)(((((b * kHF + hf1) * kGmmK + k) * T + t) * hw) + pix);
      const float ep = eps ? __ldg(eps + idx) : philox_eps(b, hf1, k, t, pix, T, hw, seed, offset);
      out1 += (e1 / sum) * (ep * expf(ls) + mu);
    }
  }
  if (V_NCHW) {
    v[(n * kHF + hf0) * hw + pix] = out0;
    if (has1) v[(n * kHF + hf1) * hw + pix] = out1;
  } else if (vpitch < 0) {      // planar latent state: quad 1 + hf/4
    v[quad_off((size_t)M, 1 + hf0 / 4, (size_t)m) + (hf0 & 3)] = out0;
    if (has1) v[quad_off((size_t)M, 1 + hf1 / 4, (size_t)m) + (hf1 & 3)] = out1;
  } else {
    v[m * vpitch + voff + hf0] = out0;
    if (has1) v[m * vpitch + voff + hf1] = out1;
  }
}

// Thread-per-pixel sampler over planar parameters (all loads coalesced): quad (j*60 + k*12 + i) holds channels
// j*240 + k*48 + 4i .. +3 of the permuted head output.
__global__ void __launch_bounds__(128) gmm_sample_planar_kernel(const float* __restrict__ params, const float* __restrict__ eps,
                                                                uint64_t seed, uint64_t offset, float* __restrict__ z, int T,
                                                                long long hw, long long M) {
  const long long m = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= M) return;
  const long long n = m / hw, pix = m - n * hw;
  const int t = (int)(n % T);
  const long long b = n / T;
  float mx[kGmmK], inv[kGmmK];
#pragma unroll
  for (int k = 0; k < kGmmK; ++k) {
    float mk = -INFINITY;
#pragma unroll
    for (int i = 0; i < 12; ++i) {
      const float4 l = __ldg(reinterpret_cast<const float4*>(params + quad_off((size_t)M, k * 12 + i, (size_t)m)));
      mk = fmaxf(fmaxf(fmaxf(mk, l.x), fmaxf(l.y, l.z)), l.w);
    }
    float sk = 0.f;
#pragma unroll
    for (int i = 0; i < 12; ++i) {
      const float4 l = __ldg(reinterpret_cast<const float4*>(params + quad_off((size_t)M, k * 12 + i, (size_t)m)));
      sk += __expf(l.x - mk) + __expf(l.y - mk) + __expf(l.z - mk) + __expf(l.w - mk);
    }
    mx[k] = mk;
    inv[k] = 1.0f / sk;
  }
  for (int i = 0; i < 12; ++i) {
    float out[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int k = 0; k < kGmmK; ++k) {
      const float4 l4 = __ldg(reinterpret_cast<const float4*>(params + quad_off((size_t)M, k * 12 + i, (size_t)m)));
      const float4 s4 = __ldg(reinterpret_cast<const float4*>(params + quad_off((size_t)M, 60 + k * 12 + i, (size_t)m)));
      const float4 m4 = __ldg(reinterpret_cast<const float4*>(params + quad_off((size_t)M, 120 + k * 12 + i, (size_t)m)));
      const float lg[4] = {l4.x, l4.y, l4.z, l4.w}, ls[4] = {s4.x, s4.y, s4.z, s4.w}, mu[4] = {m4.x, m4.y, m4.z, m4.w};
      float ep4[4];
      if (eps) {
#pragma unroll
        for (int e = 0; e < 4; ++e) ep4[e] = __ldg(eps + (uint64_t)(((((b * kHF + 4 * i + e) * kGmmK + k) * T + t) * hw) + pix));
      } else {
        philox_normal4(eps_group(b, i, k, t, pix, T, hw), seed, offset, ep4);   // one Philox call per (k, hf quad)
      }
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float ep = ep4[e];
        // ex2.approx based exp: relative error ~2^-21 on arguments in [-7, 7] / (-inf, 0]
        const float pi = __expf(lg[e] - mx[k]) * inv[k];
        out[e] += pi * (ep * __expf(fminf(fmaxf(ls[e], -7.f), 7.f)) + mu[e]);
      }
    }
    store4(z + quad_off((size_t)M, 1 + i, (size_t)m), make_float4(out[0], out[1], out[2], out[3]));
  }
}

// Warp-split form of the planar sampler (default): a CTA of four warps owns 32 pixels and warp w owns HF quads 3w .. 3w+2, so
// the 720 parameters of a pixel are read from DRAM exactly once (the thread-per-pixel form above walks the 240 logits three
// times and its third walk misses L1/L2: 3.2 GB read per 1080p GOP for 2.6 GB of parameters).  The softmax over the 48 HF
// channels is independent per mixture component, so the CTA walks k = 0..4 and a thread only holds the 12 logits of (its three
// quads, this k) plus 12 accumulators: 64 registers, 32 resident warps per SM -- the kernel is issue-bound (60 Philox calls,
// 120 Box-Muller pairs and 480 exp per pixel), so occupancy is what the form with all 60 logits in registers (113 registers,
// 16 warps per SM: slower than thread-per-pixel in the stream) lacked.  The per-component max and exp-sum are combined across
// the four warps through shared memory in a fixed order (deterministic).
// PT = float, or __half: the tcgen05 head stores the parameters as fp16 quads (8 bytes) -- its inputs are bf16 (2^-9), so fp16's
// 2^-11 adds nothing measurable, and the 720-channel tensor's write + read halve (2.6 -> 1.3 GB each per 1080p GOP)
template <typename PT>
__device__ __forceinline__ float4 load_quad(const PT* params, size_t quad_index) {
  if constexpr (std::is_same<PT, __half>::value) {
    const uint2 r = __ldg(reinterpret_cast<const uint2*>(params + quad_index));
    const __half2 a = *reinterpret_cast<const __half2*>(&r.x), b = *reinterpret_cast<const __half2*>(&r.y);
    const float2 fa = __half22float2(a), fb = __half22float2(b);
    return make_float4(fa.x, fa.y, fb.x, fb.y);
  } else {
    return __ldg(reinterpret_cast<const float4*>(params + quad_index));
  }
}

template <bool kEps, typename PT>
__global__ void __launch_bounds__(128, 8) gmm_sample_planar_perk_kernel(const PT* __restrict__ params, const float* __restrict__ eps,
                                                                        uint64_t seed, uint64_t offset, float* __restrict__ z, int T,
                                                                        long long hw, long long M) {
  __shared__ float red_max[4][32];
  __shared__ float red_sum[4][32];
  const int lane = threadIdx.x & 31, wq = threadIdx.x >> 5;
  const long long m_raw = (long long)blockIdx.x * 32 + lane;
  const bool live = m_raw < M;
  const long long m = live ? m_raw : M - 1;          // out-of-range lanes shadow the last pixel (they must reach the barriers)
  const long long n = m / hw, pix = m - n * hw;
  const int t = (int)(n % T);
  const long long b = n / T;
  float out[3][4];
#pragma unroll
  for (int q = 0; q < 3; ++q)
#pragma unroll
    for (int c = 0; c < 4; ++c) out[q][c] = 0.f;
#pragma unroll 1
  for (int k = 0; k < kGmmK; ++k) {
    float e[3][4];
#pragma unroll
    for (int q = 0; q < 3; ++q) {
      const int i = wq * 3 + q;
      const float4 l = load_quad<PT>(params, quad_off((size_t)M, k * 12 + i, (size_t)m));
      e[q][0] = l.x; e[q][1] = l.y; e[q][2] = l.z; e[q][3] = l.w;
    }
    float mk = e[0][0];
#pragma unroll
    for (int q = 0; q < 3; ++q)
#pragma unroll
      for (int c = 0; c < 4; ++c) mk = fmaxf(mk, e[q][c]);
    red_max[wq][lane] = mk;
    __syncthreads();
    mk = fmaxf(fmaxf(red_max[0][lane], red_max[1][lane]), fmaxf(red_max[2][lane], red_max[3][lane]));
    float sk = 0.f;
#pragma unroll
    for (int q = 0; q < 3; ++q)
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        e[q][c] = __expf(e[q][c] - mk);              // ex2.approx based exp: relative error ~2^-21 on (-inf, 0]
        sk += e[q][c];
      }
    red_sum[wq][lane] = sk;
    __syncthreads();
    const float inv = 1.0f / (((red_sum[0][lane] + red_sum[1][lane]) + red_sum[2][lane]) + red_sum[3][lane]);
#pragma unroll
    for (int q = 0; q < 3; ++q) {
      const int i = wq * 3 + q;
      // issued before the Philox rounds, which cover most of their latency
      const float4 s4 = load_quad<PT>(params, quad_off((size_t)M, 60 + k * 12 + i, (size_t)m));
      const float4 m4 = load_quad<PT>(params, quad_off((size_t)M, 120 + k * 12 + i, (size_t)m));
      const float ls[4] = {s4.x, s4.y, s4.z, s4.w}, mu[4] = {m4.x, m4.y, m4.z, m4.w};
      float ep4[4];
      if constexpr (kEps) {
#pragma unroll
        for (int c = 0; c < 4; ++c) ep4[c] = __ldg(eps + (uint64_t)(((((b * kHF + 4 * i + c) * kGmmK + k) * T + t) * hw) + pix));
      } else {
        philox_normal4(eps_group(b, i, k, t, pix, T, hw), seed, offset, ep4);   // one Philox call per (k, hf quad)
      }
#pragma unroll
      for (int c = 0; c < 4; ++c)
        out[q][c] += (e[q][c] * inv) * (ep4[c] * __expf(fminf(fmaxf(ls[c], -7.f), 7.f)) + mu[c]);
    }
  }
  if (live) {
#pragma unroll
    for (int q = 0; q < 3; ++q)
      store4(z + quad_off((size_t)M, 1 + wq * 3 + q, (size_t)m), make_float4(out[q][0], out[q][1], out[q][2], out[q][3]));
  }
}

// tail_gmm.5 rows hf*15+k*3+j  ->  j*240 + k*48 + hf  (so each tcgen05 pass emits one parameter kind, hf contiguous)
__global__ void permute_gmm_rows_kernel(const float* __restrict__ w, const float* __restrict__ b, float* __restrict__ wp,
                                        float* __restrict__ bp, int by_component) {
  const int np = blockIdx.x;                      // permuted row
  const int hf = np % 48;
  const int j = by_component ? (np % 144) / 48 : np / 240;
  const int k = by_component ? np / 144 : (np % 240) / 48;
  const int src = hf * 15 + k * 3 + j;
  for (int c = threadIdx.x; c < 256; c += blockDim.x) wp[np * 256 + c] = w[src * 256 + c];
  if (threadIdx.x == 0) bp[np] = b[src];
}

// eps [B,48,5,T,h,w]: thread i of [B,12,5,T,hw] writes the four hf of its quad
__global__ void export_eps_kernel(float* __restrict__ eps, uint64_t seed, uint64_t offset, int T, long long hw, long long ngroups) {
  const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= ngroups) return;
  float n4[4];
  philox_normal4((uint64_t)g, seed, offset, n4);
  const long long inner = (long long)kGmmK * T * hw;          // elements per hf channel of one clip
  const long long bq = g / inner, rest = g - bq * inner;      // bq = b*12 + hf/4
  const long long b = bq / (kHF / 4);
  const int hfq = (int)(bq - b * (kHF / 4));
#pragma unroll
  for (int e = 0; e < 4; ++e) eps[(b * kHF + hfq * 4 + e) * inner + rest] = n4[e];
}

// ---- launchers ----------------------------------------------------------------------------------------
int launch_ga_wmap(const float* fcw, float* wmap, int h, int w, cudaStream_t st) {
  ga_wmap_kernel<<<cdiv((long long)h * w, 128), 128, 0, st>>>(fcw, wmap, h, w);
  SELFC_LAUNCH_CHECK("ga_wmap_kernel");
  return 0;
}

template <typename T>
int launch_ga_stat(const T* x, int pitch, const float* wmap, float* partial, int nsplit, int BT, int hw, cudaStream_t st) {
  ga_stat_kernel<T><<<dim3(nsplit, BT), 256, 0, st>>>(x, pitch, wmap, partial, nsplit, hw);
  SELFC_LAUNCH_CHECK("ga_stat_kernel");
  return 0;
}
template int launch_ga_stat<float>(const float*, int, const float*, float*, int, int, int, cudaStream_t);
template int launch_ga_stat<__nv_bfloat16>(const __nv_bfloat16*, int, const float*, float*, int, int, int, cudaStream_t);
template int launch_ga_stat<bfx2>(const bfx2*, int, const float*, float*, int, int, int, cudaStream_t);

int launch_ga_weights(const float* partial, int nsplit, const float* fcb, const float* p2w, const float* p2b, const float* p3w,
                      const float* p3b, float* wmat, float* wsum, int B, int T, cudaStream_t st) {
  SELFC_CHECK_ARG(T >= 1 && T <= 32, "GlobalAgg: temporal length %d outside [1,32]", T);
  const size_t smem = (size_t)(T * 64 + 2 * T * 65 + T * 32 + 16 * 64 + 2 * 64 * kGaWPitch) * sizeof(float);   // 41 KB at T = 7, 64 KB at T = 32
  static bool smem_set[64] = {};           // per device: function attributes belong to the device's context
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev >= 0 && dev < 64 && !smem_set[dev]) {
    SELFC_CUDA(cudaFuncSetAttribute(ga_weights_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 72 * 1024));
    smem_set[dev] = true;
  }
  ga_weights_kernel<<<B, 1024, smem, st>>>(partial, nsplit, fcb, p2w, p2b, p3w, p3b, wmat, wsum, T);
  SELFC_LAUNCH_CHECK("ga_weights_kernel");
  return 0;
}

int launch_gmm_sample(const float* params, bool params_nchw, const float* eps, uint64_t seed, uint64_t offset, float* v,
                      bool v_nchw, int vpitch, int voff, int B, int T, int h, int w, cudaStream_t st) {
  const long long M = (long long)B * T * h * w;
  if (M == 0) return 0;
  const int grid = cdiv(M, 4);
  if (params_nchw && v_nchw)
    gmm_sample_kernel<true, true><<<grid, 128, 0, st>>>(params, eps, seed, offset, v, vpitch, voff, T, h, w, M);
  else if (!params_nchw && !v_nchw)
    gmm_sample_kernel<false, false><<<grid, 128, 0, st>>>(params, eps, seed, offset, v, vpitch, voff, T, h, w, M);
  else if (!params_nchw && v_nchw)
    gmm_sample_kernel<false, true><<<grid, 128, 0, st>>>(params, eps, seed, offset, v, vpitch, voff, T, h, w, M);
  else
    gmm_sample_kernel<true, false><<<grid, 128, 0, st>>>(params, eps, seed, offset, v, vpitch, voff, T, h, w, M);
  SELFC_LAUNCH_CHECK("gmm_sample_kernel");
  return 0;
}

int launch_gmm_sample_planar(const float* params, const float* eps, uint64_t seed, uint64_t offset, float* z, int B, int T, int h,
                             int w, cudaStream_t st, int form, bool params_half) {
  const long long M = (long long)B * T * h * w;
  if (M == 0) return 0;
  static int split = -1;                  // SELFC_GMM_SPLIT=0: the thread-per-pixel form
  if (split < 0) {
    const char* e = getenv("SELFC_GMM_SPLIT");
    split = (e && atoi(e) == 0) ? 0 : 1;
  }
  const int f = form < 0 ? split : form;
  if (params_half) {
    SELFC_CHECK_ARG(f == 1, "gmm_sample_planar: fp16 parameters are read by the per-component kernel only");
    const __half* ph = reinterpret_cast<const __half*>(params);
    if (eps != nullptr) gmm_sample_planar_perk_kernel<true, __half><<<cdiv(M, 32), 128, 0, st>>>(ph, eps, seed, offset, z, T, (long long)h * w, M);
    else gmm_sample_planar_perk_kernel<false, __half><<<cdiv(M, 32), 128, 0, st>>>(ph, eps, seed, offset, z, T, (long long)h * w, M);
  } else if (f == 1 && eps != nullptr)
    gmm_sample_planar_perk_kernel<true, float><<<cdiv(M, 32), 128, 0, st>>>(params, eps, seed, offset, z, T, (long long)h * w, M);
  else if (f == 1)
    gmm_sample_planar_perk_kernel<false, float><<<cdiv(M, 32), 128, 0, st>>>(params, eps, seed, offset, z, T, (long long)h * w, M);
  else
    gmm_sample_planar_kernel<<<cdiv(M, 128), 128, 0, st>>>(params, eps, seed, offset, z, T, (long long)h * w, M);
  SELFC_LAUNCH_CHECK("gmm_sample_planar_kernel");
  return 0;
}

int launch_permute_gmm_rows(const float* w, const float* b, float* wp, float* bp, bool by_component, cudaStream_t st) {
  permute_gmm_rows_kernel<<<720, 128, 0, st>>>(w, b, wp, bp, by_component ? 1 : 0);
  SELFC_LAUNCH_CHECK("permute_gmm_rows_kernel");
  return 0;
}

int launch_export_eps(float* eps, uint64_t seed, uint64_t offset, int B, int T, long long hw, cudaStream_t st) {
  const long long n = (long long)B * (kHF / 4) * kGmmK * T * hw;
  if (n == 0) return 0;
  export_eps_kernel<<<cdiv(n, 256), 256, 0, st>>>(eps, seed, offset, T, hw, n);
  SELFC_LAUNCH_CHECK("export_eps_kernel");
  return 0;
}

}  // namespace selfc
