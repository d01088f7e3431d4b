// (3,1,1) temporal convolution of D2DTInput (conv5) with the InvBlockExp coupling fused into its epilogue, and the
// GlobalAgg apply step, as one tcgen05 / TMEM / TMA kernel (BF16 mode).
//
// Reference behaviour restated (not copied):
//   conv5                      Subnet_constructor.py:106,130     out[t] = sum_{dt} W[dt] . in[t+dt-1], zero-padded at clip ends
//   coupling                   SelfC_GMM_arch_inv.py:21-33      y1 = x1 +/- F, s = 2*sigmoid(H)-1, y2 = x2*e^s + G | (x2-G)/e^s
//   GlobalAgg apply            SelfC_GMM_arch_inv.py:266,278-285 out[t'] = x[t'] + sum_t W[b,t,t'] * proj1(x[t])
//
// Mapping: a CTA owns 128 consecutive pixels of ALL T frames of one clip.  Per 16-channel K slice the TMA brings the
// T frame tiles ([2 planes][128 px][8 ch], the no-swizzle K-major core-matrix layout) into one pipeline stage ONCE;
// frame t's tile is the A operand of up to three MMAs (output frames t-1, t, t+1), so the temporal taps re-use the
// load.  Each output frame has its own fp32 accumulator (N columns) in TMEM: T*N <= 512 columns.  Clip-end zero
// padding = the corresponding MMA is simply not issued.  Weights stay resident in shared memory.
// Epilogue warps read the accumulators frame by frame (tcgen05.ld), apply bias and the fused coupling / residual
// arithmetic on the fp32 latent state, and release each frame's accumulator as soon as it is drained so the next
// tile's MMAs chase the epilogue.
#include <cuda.h>

#include "common.cuh"
#include "conv_tc.h"
#include "kernels.h"
#include "tc_ptx.cuh"

namespace selfc {
namespace tc5 {

using namespace tc;

constexpr int MT = 128;
constexpr int FRAME_BYTES = 2 * MT * 16;   // one frame tile of one 16-channel slice
constexpr int NST = 4;
constexpr int TMAX = 8;
constexpr int THREADS = 192;
constexpr int BAR_BYTES = 512;
constexpr int BIAS_BYTES = 1024;   // up to 256 fp32

struct Params {
  const void* wimg;
  const float* bias;
  int T, B, hw, nchunk, npad, taps, cout;
  int tiles_p, ntiles, tmem_cols;
  long long m_limit;   // rows (pixels) that really exist in the buffer
  int epi, rev, act;
  __nv_bfloat16* outT;
  int outT_pitch, outT_off;
  float* outF;
  int outF_pitch, outF_off;
  float* z;
  float* sbuf;
  __nv_bfloat16* copyA;
  int copyA_pitch;
  __nv_bfloat16* copyB;
  int copyB_pitch;
  int copy_pad;
  const float* wmat;
  const float* wsum;
  const __nv_bfloat16* resid;
  int resid_pitch;
  __nv_bfloat16* outAct;
  int outAct_pitch;
  int* err;
};

__device__ __forceinline__ uint32_t pack_bf2(float a, float b) {
  __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ void store_bf16x8(__nv_bfloat16* o, const float* v) {
  uint4 pk;
  pk.x = pack_bf2(v[0], v[1]); pk.y = pack_bf2(v[2], v[3]); pk.z = pack_bf2(v[4], v[5]); pk.w = pack_bf2(v[6], v[7]);
  *reinterpret_cast<uint4*>(o) = pk;
}
__device__ __forceinline__ void load_bf16x8(const __nv_bfloat16* p, float* v) {
  const uint4 r = *reinterpret_cast<const uint4*>(p);
  const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    __nv_bfloat162 b = *reinterpret_cast<const __nv_bfloat162*>(&w[i]);
    v[2 * i] = __low2float(b);
    v[2 * i + 1] = __high2float(b);
  }
}
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t r[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr)
               : "memory");
}

__global__ void __launch_bounds__(THREADS, 1) temporal_tc_kernel(const __grid_constant__ CUtensorMap tmap, const Params p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gen_base = smem_raw + (base - smem_u32(smem_raw));
  const int T = p.T;
  const uint32_t stage_bytes = (uint32_t)T * FRAME_BYTES;
  const uint32_t a_base = base;
  const uint32_t bar_base = base + NST * stage_bytes;
  const uint32_t bias_off = NST * stage_bytes + BAR_BYTES;
  const uint32_t w_base = base + bias_off + BIAS_BYTES;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (NST + s); };
  const uint32_t w_bar = bar_base + 8u * (2 * NST);
  const uint32_t tfull_bar = bar_base + 8u * (2 * NST + 1);
  auto tempty_bar = [&](int t) { return bar_base + 8u * (2 * NST + 2 + t); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * NST + 2 + TMAX);
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(gen_base + NST * stage_bytes + 8 * (2 * NST + 2 + TMAX));
  float* sbias = reinterpret_cast<float*>(gen_base + bias_off);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < NST; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    mbar_init(w_bar, 1);
    mbar_init(tfull_bar, 1);
    for (int t = 0; t < TMAX; ++t) mbar_init(tempty_bar(t), 4);
    fence_barrier_init();
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap) : "memory");
  }
  if (warp == 1) tmem_alloc(tmem_slot, (uint32_t)p.tmem_cols);
  for (int i = threadIdx.x; i < p.npad; i += THREADS) sbias[i] = __ldg(p.bias + i);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  const int nchunk = p.nchunk, npad = p.npad, taps = p.taps;
  const uint32_t wtile = (uint32_t)npad * 32u;

  if (warp == 0) {
    if (lane == 0) {
      // ===================== TMA producer =====================
      const uint32_t wbytes = (uint32_t)taps * nchunk * wtile;
      mbar_expect_tx(w_bar, wbytes);
      for (int tap = 0; tap < taps; ++tap)
        bulk_g2s(w_base + tap * nchunk * wtile, (const uint8_t*)p.wimg + (size_t)tap * nchunk * wtile, (uint32_t)nchunk * wtile, w_bar);
      int s = 0;
      uint32_t ph = 0;
      for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x) {
        const int b = tile / p.tiles_p;
        const int p0 = (tile - b * p.tiles_p) * MT;
        for (int c = 0; c < nchunk; ++c) {
          mbar_wait(empty_bar(s), ph ^ 1u, p.err, 11);
          mbar_expect_tx(full_bar(s), stage_bytes);
          const uint32_t dst = a_base + s * stage_bytes;
          for (int t = 0; t < T; ++t) {
            tma_load_3d(dst + t * FRAME_BYTES, &tmap, full_bar(s), c * 16, p0, b * T + t);
            tma_load_3d(dst + t * FRAME_BYTES + MT * 16, &tmap, full_bar(s), c * 16 + 8, p0, b * T + t);
          }
          if (++s == NST) { s = 0; ph ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // ===================== MMA issuer =====================
      const uint32_t idesc = umma_idesc_bf16(128, npad);
      mbar_wait(w_bar, 0, p.err, 12);
      int s = 0;
      uint32_t ph = 0;
      int it = 0;
      for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x, ++it) {
        for (int c = 0; c < nchunk; ++c) {
          mbar_wait(full_bar(s), ph, p.err, 14);
          tc_fence_after();
          const uint32_t a_stage = a_base + s * stage_bytes;
          for (int to = 0; to < T; ++to) {
            if (c == 0) {
              mbar_wait(tempty_bar(to), ((uint32_t)it & 1u) ^ 1u, p.err, 13);
              tc_fence_after();
            }
            bool first = (c == 0);
            for (int tap = 0; tap < taps; ++tap) {
              const int ti = taps == 3 ? to + tap - 1 : to;
              if (ti < 0 || ti >= T) continue;
              const uint64_t ad = umma_desc(a_stage + (uint32_t)ti * FRAME_BYTES, MT * 16, 128);
              const uint64_t bd = umma_desc(w_base + (uint32_t)(tap * nchunk + c) * wtile, (uint32_t)npad * 16u, 128);
              umma_bf16(tmem_base + (uint32_t)(to * npad), ad, bd, idesc, first ? 0u : 1u);
              first = false;
            }
          }
          umma_commit(empty_bar(s));
          if (++s == NST) { s = 0; ph ^= 1u; }
        }
        umma_commit(tfull_bar);
      }
    }
  } else {
    // ===================== epilogue warps 2..5 =====================
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16);
    int it = 0;
    for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x, ++it) {
      const int b = tile / p.tiles_p;
      const int pix = (tile - b * p.tiles_p) * MT + row;
      const bool valid = pix < p.hw;
      // the epilogue's operands (latent state, log-scale, residual) are pulled into L2 while the MMAs of this tile run
      if (pix < p.hw) {
        for (int t = 0; t < T; ++t) {
          const size_t m = ((size_t)b * T + t) * p.hw + pix;
          if (p.epi == EPI_COUPLE_Y1) {
            prefetch_l2(p.z + m * kZPitch);
          } else if (p.epi == EPI_COUPLE_Y2) {
            prefetch_l2(p.z + m * kZPitch);
            prefetch_l2(p.z + m * kZPitch + 32);
            prefetch_l2(p.sbuf + m * kHF);
            prefetch_l2(p.sbuf + m * kHF + 32);
          } else if (p.epi == EPI_GA) {
            prefetch_l2(p.resid + m * p.resid_pitch);
          }
        }
      }
      mbar_wait(tfull_bar, (uint32_t)it & 1u, p.err, 15);
      tc_fence_after();
      if (p.epi == EPI_GA) {
        // out[t'] = x[t'] + bias * colsum(W)[t'] + sum_t W[b,t,t'] * D_t
        const float* wm = p.wmat + (size_t)b * T * T;
        for (int n0 = 0; n0 < p.cout; n0 += 8) {
          uint32_t d[TMAX][8];
#pragma unroll
          for (int t = 0; t < TMAX; ++t)
            if (t < T) tmem_ld8(lane_addr + (uint32_t)(t * npad + n0), d[t]);
          tmem_ld_wait();
          if (valid) {
            for (int tp = 0; tp < T; ++tp) {
              const size_t m = ((size_t)b * T + tp) * p.hw + pix;
              float v[8];
              load_bf16x8(p.resid + m * p.resid_pitch + n0, v);
              const float ws = __ldg(p.wsum + b * T + tp);
#pragma unroll
              for (int j = 0; j < 8; ++j) v[j] += sbias[n0 + j] * ws;
#pragma unroll
              for (int t = 0; t < TMAX; ++t) {
                if (t < T) {
                  const float wv = __ldg(wm + t * T + tp);
#pragma unroll
                  for (int j = 0; j < 8; ++j) v[j] = fmaf(wv, __uint_as_float(d[t][j]), v[j]);
                }
              }
              if (p.outT) store_bf16x8(p.outT + m * p.outT_pitch + p.outT_off + n0, v);
              if (p.outF) {
                float* o = p.outF + m * p.outF_pitch + n0;
                store4(o, make_float4(v[0], v[1], v[2], v[3]));
                store4(o + 4, make_float4(v[4], v[5], v[6], v[7]));
              }
              if (p.outAct) {
#pragma unroll
                for (int j = 0; j < 8; ++j) v[j] = lrelu02(v[j]);
                store_bf16x8(p.outAct + m * p.outAct_pitch + n0, v);
              }
            }
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0)
          for (int t = 0; t < T; ++t) mbar_arrive(tempty_bar(t));
        continue;
      }
      for (int t = 0; t < T; ++t) {
        const size_t m = ((size_t)b * T + t) * p.hw + pix;
        const int ncols = p.epi == EPI_COUPLE_Y1 ? 16 : (p.epi == EPI_STORE ? ((p.cout + 15) & ~15) : kHF);
        for (int n0 = 0; n0 < ncols; n0 += 16) {
          uint32_t r[16];
          tmem_ld16(lane_addr + (uint32_t)(t * npad + n0), r);
          tmem_ld_wait();
          if (!valid || (long long)m >= p.m_limit) continue;
          float v[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(r[j]) + sbias[n0 + j];
          switch (p.epi) {
            case EPI_STORE: {
              if (p.act) {
#pragma unroll
                for (int j = 0; j < 16; ++j) v[j] = lrelu02(v[j]);
              }
              if (p.outT) {
                __nv_bfloat16* o = p.outT + m * p.outT_pitch + p.outT_off + n0;
                if (n0 + 16 <= p.cout) {
                  store_bf16x8(o, v);
                  store_bf16x8(o + 8, v + 8);
                } else {
                  for (int j = 0; j < 16; ++j)
                    if (n0 + j < p.cout) o[j] = __float2bfloat16_rn(v[j]);
                }
              }
              if (p.outF) {
                float* o = p.outF + m * p.outF_pitch + p.outF_off + n0;
                if (n0 + 16 <= p.cout && ((p.outF_pitch | p.outF_off) & 3) == 0) {
#pragma unroll
                  for (int j = 0; j < 16; j += 4) store4(o + j, make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]));
                } else {
                  for (int j = 0; j < 16; ++j)
                    if (n0 + j < p.cout) o[j] = v[j];
                }
              }
            } break;
            case EPI_COUPLE_Y1: {
              float* zp = p.z + m * kZPitch;
              const float4 x1 = load4(zp);
              float y[16];
#pragma unroll
              for (int j = 0; j < 16; ++j) y[j] = 0.f;
              y[0] = p.rev ? x1.x - v[0] : x1.x + v[0];
              y[1] = p.rev ? x1.y - v[1] : x1.y + v[1];
              y[2] = p.rev ? x1.z - v[2] : x1.z + v[2];
              store4(zp, make_float4(y[0], y[1], y[2], 0.f));
              if (p.copyA) {
                store_bf16x8(p.copyA + m * p.copyA_pitch, y);
                if (p.copy_pad > 8) store_bf16x8(p.copyA + m * p.copyA_pitch + 8, y + 8);
              }
              if (p.copyB) {
                store_bf16x8(p.copyB + m * p.copyB_pitch, y);
                if (p.copy_pad > 8) store_bf16x8(p.copyB + m * p.copyB_pitch + 8, y + 8);
              }
            } break;
            case EPI_COUPLE_S: {
              float* sp = p.sbuf + m * kHF + n0;
#pragma unroll
              for (int j = 0; j < 16; j += 4) {
                float s[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) s[e] = (1.0f / (1.0f + expf(-v[j + e]))) * 2.0f - 1.0f;
                store4(sp + j, make_float4(s[0], s[1], s[2], s[3]));
              }
            } break;
            case EPI_COUPLE_Y2: {
              float* zp = p.z + m * kZPitch + kZHf + n0;
              const float* sp = p.sbuf + m * kHF + n0;
              float y[16];
#pragma unroll
              for (int j = 0; j < 16; j += 4) {
                const float4 x2 = load4(zp + j);
                const float4 s4 = load4(sp + j);
                const float xr[4] = {x2.x, x2.y, x2.z, x2.w};
                const float sr[4] = {s4.x, s4.y, s4.z, s4.w};
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  const float ex = expf(sr[e]);
                  y[j + e] = p.rev ? (xr[e] - v[j + e]) / ex : xr[e] * ex + v[j + e];
                }
                store4(zp + j, make_float4(y[j], y[j + 1], y[j + 2], y[j + 3]));
              }
              if (p.copyA) {
                store_bf16x8(p.copyA + m * p.copyA_pitch + n0, y);
                store_bf16x8(p.copyA + m * p.copyA_pitch + n0 + 8, y + 8);
              }
            } break;
            default: break;
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(tempty_bar(t));
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
  }
}

// wref [cout][cin_ref][taps] fp32 -> bf16 image [tap][kstep][kcore(2)][ngroup(npad/8)][n%8][k%8]
__global__ void pack_temporal_kernel(const float* __restrict__ wref, const float* __restrict__ bref, __nv_bfloat16* __restrict__ img,
                                     float* __restrict__ bias, int cout, int cin_ref, int taps, int cin_buf, int xreal, int xpad,
                                     int npad) {
  const int nchunk = cin_buf / 16;
  const int total = taps * cin_buf * npad;
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx < npad) bias[idx] = idx < cout ? bref[idx] : 0.f;
  if (idx >= total) return;
  const int n = idx % npad;
  const int c = (idx / npad) % cin_buf;
  const int tap = idx / (npad * cin_buf);
  const int cref = c < xreal ? c : (c < xpad ? -1 : c - xpad + xreal);
  float v = 0.f;
  if (n < cout && cref >= 0 && cref < cin_ref) v = wref[((size_t)n * cin_ref + cref) * taps + tap];
  const int ks = c / 16, kk = c % 16;
  const size_t off = (size_t)(tap * nchunk + ks) * (npad * 16) + (size_t)((kk / 8) * (npad / 8) + n / 8) * 64 + (n % 8) * 8 + (kk % 8);
  img[off] = __float2bfloat16_rn(v);
}

}  // namespace tc5

int pack_temporal_weights(TcTempW& w, const float* wref, const float* bref, int cout, int cin_ref, int taps, int cin_buf, int xreal,
                          int xpad, cudaStream_t st) {
  SELFC_CHECK_ARG(cin_buf % 16 == 0 && cout >= 1 && cout <= 256, "temporal_tc: cin %d / cout %d unsupported", cin_buf, cout);
  const int npad = (cout + 15) & ~15;
  const size_t bytes = (size_t)taps * (cin_buf / 16) * npad * 32;
  if (w.img == nullptr || w.img_bytes != bytes) {
    free_temporal_weights(w);
    SELFC_CUDA(cudaMalloc(&w.img, bytes));
    SELFC_CUDA(cudaMalloc(&w.bias, 256 * sizeof(float)));
    w.img_bytes = bytes;
  }
  w.cin_buf = cin_buf; w.npad = npad; w.taps = taps; w.cout = cout;
  const int total = taps * cin_buf * npad;
  tc5::pack_temporal_kernel<<<cdiv(total, 256), 256, 0, st>>>(wref, bref, reinterpret_cast<__nv_bfloat16*>(w.img), w.bias, cout, cin_ref,
                                                              taps, cin_buf, xreal, xpad, npad);
  SELFC_LAUNCH_CHECK("pack_temporal_kernel");
  return 0;
}

void free_temporal_weights(TcTempW& w) {
  if (w.img) cudaFree(w.img);
  if (w.bias) cudaFree(w.bias);
  w.img = nullptr;
  w.bias = nullptr;
  w.img_bytes = 0;
}

bool temporal_tc_supported(const TcTempW& w, int T) { return w.img != nullptr && T >= 1 && T <= tc5::TMAX && T * w.npad <= 512; }

int launch_temporal_tc(const TcTempW& w, const TcTempArgs& a, cudaStream_t st) {
  SELFC_CHECK_ARG(temporal_tc_supported(w, a.T), "temporal_tc: T=%d x N=%d does not fit TMEM", a.T, w.npad);
  SELFC_CHECK_ARG(a.in_pitch % 8 == 0 && aligned16(a.in), "temporal_tc: input pitch/alignment");
  tc::EncodeTiledFn encode = tc::get_encode_fn();
  if (!encode) {
    set_error("cuTensorMapEncodeTiled is not available from this driver");
    return SELFC_E_CUDA;
  }
  const int BT = a.B * a.T;
  SELFC_CHECK_ARG(a.epi != EPI_GA || w.npad <= 64, "temporal_tc: GlobalAgg epilogue needs N <= 64");
  CUtensorMap tmap;
  const cuuint64_t gdim[3] = {(cuuint64_t)a.in_pitch, (cuuint64_t)a.hw, (cuuint64_t)BT};
  const cuuint64_t gstr[2] = {(cuuint64_t)a.in_pitch * 2, (cuuint64_t)a.hw * a.in_pitch * 2};
  const cuuint32_t box[3] = {8, (cuuint32_t)tc5::MT, 1};
  const cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = encode(&tmap, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<__nv_bfloat16*>(a.in), gdim, gstr, box, estr,
                      CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled (temporal) failed with CUresult %d", (int)r);
    return SELFC_E_CUDA;
  }
  tc5::Params p;
  memset(&p, 0, sizeof(p));
  p.wimg = w.img; p.bias = w.bias;
  p.T = a.T; p.B = a.B; p.hw = a.hw; p.nchunk = w.cin_buf / 16; p.npad = w.npad; p.taps = w.taps; p.cout = w.cout;
  p.tiles_p = cdiv(a.hw, tc5::MT);
  p.ntiles = p.tiles_p * a.B;
  int cols = a.T * w.npad, pw = 32;
  while (pw < cols) pw <<= 1;
  p.tmem_cols = pw;
  p.epi = a.epi; p.rev = a.rev; p.act = a.act;
  p.outT = a.outT; p.outT_pitch = a.outT_pitch; p.outT_off = a.outT_off;
  p.outF = a.outF; p.outF_pitch = a.outF_pitch; p.outF_off = a.outF_off;
  p.m_limit = a.m_limit > 0 ? a.m_limit : (long long)BT * a.hw;
  p.z = a.z; p.sbuf = a.sbuf;
  p.copyA = a.copyA; p.copyA_pitch = a.copyA_pitch; p.copyB = a.copyB; p.copyB_pitch = a.copyB_pitch; p.copy_pad = a.copy_pad;
  p.wmat = a.wmat; p.wsum = a.wsum; p.resid = a.resid; p.resid_pitch = a.resid_pitch;
  p.outAct = a.outAct; p.outAct_pitch = a.outAct_pitch;
  p.err = tc::err_flag_for_device();
  if (p.ntiles == 0) return 0;
  const int smem = tc5::NST * a.T * tc5::FRAME_BYTES + tc5::BAR_BYTES + tc5::BIAS_BYTES + (int)w.img_bytes + 1024;
  static int smem_set = 0;
  if (smem_set < smem) {
    SELFC_CUDA(cudaFuncSetAttribute(tc5::temporal_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    smem_set = 227 * 1024;
  }
  SELFC_CHECK_ARG(smem <= 227 * 1024, "temporal_tc: %d bytes of shared memory needed", smem);
  const int nsm = tc::num_sms();
  const int grid = p.ntiles < nsm ? p.ntiles : nsm;
  tc5::temporal_tc_kernel<<<grid, tc5::THREADS, smem, st>>>(tmap, p);
  SELFC_LAUNCH_CHECK("temporal_tc_kernel");
  return 0;
}

}  // namespace selfc
