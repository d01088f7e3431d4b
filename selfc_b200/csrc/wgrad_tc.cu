// Weight gradients of the dense-block convolutions on the tensor cores (training step, BF16X3 mode; SURVEY row a13:
// models/SelfC_model.py:148-183 -- what autograd's conv3d weight-gradient kernels compute for Subnet_constructor.py:102-106).
//
//   dW[tap][c][n] = sum over pixels p of  in[p + shift(tap)][c] * g[p][n]        (zero outside the frame / the clip)
//
// is a GEMM with the PIXELS as the contraction dimension.  Both operands are first transposed into zero-padded pixel PLANES
//   AT[row][P]   row 0 = 1 at every real pixel (its products are the bias gradient), row 1 + c = activation channel c
//   GT[n][P]     the (LeakyReLU-masked) output gradient
// as (hi, lo) bf16 pairs, P = ((f' * (h+2) + y+1) * Wp + x+1) with one zero row above and below every frame, zero columns up to
// Wp = roundup(w + 2, 8) and one zero frame before every clip and after the last one.  In these planes EVERY tap is a plain offset
// of the pixel index -- (ky-1) * Wp + (kx-1) for a spatial tap, (dt-1) * (h+2) * Wp for a temporal one -- and every out-of-frame /
// out-of-clip neighbour is a stored zero, so a tap's partial sum is one K-major GEMM over a range of P with the A operand loaded
// at a shifted coordinate (TMA, out-of-range coordinates zero-filled):
//   D[sh][row][kx * 32 + n] += sum_P AT[row][P + (sh-1) * sh_stride] * GT_kx[n][P],  GT_kx[n][P] = g[n][P - (kx-1)]
// (spatial: sh = ky, and the three kx taps are stacked in N as three copies of the gradient plane written one pixel apart: a TMA
// box must start on a 16-byte boundary of its innermost dimension, so the one-pixel shifts cannot be load coordinates; the row
// shifts are multiples of Wp = 8 k pixels and can).
// M = 128 rows, N = 96 (3 x 32) / the padded output count of conv5, K = 16 pixels per MMA, three MMAs per product (hi.hi, hi.lo,
// lo.hi), fp32 accumulation in tensor memory over the CTA's share of the pixels (split-K), then one vector reduction (red.v4.f32)
// per four weights into the fp32 scratch the existing unpack kernel reads.
#include <cuda.h>
#include <stdlib.h>
#include <string.h>

#include "net_ctx.h"
#include "tc_ptx.cuh"
#include "wgrad_tc.h"

namespace selfc {
namespace wg {

using namespace tc;

constexpr int KT = 32;                  // pixels per pipeline stage: one 64-byte row of bf16 per operand row (SWIZZLE_64B)
constexpr int ROWS = 128;               // AT rows per CTA (MMA M)
constexpr int A_TILE = ROWS * KT * 2;   // 8 KB
constexpr int THREADS = 192;
constexpr int NST_MAX = 4;
constexpr int BAR_BYTES = 256;

struct Params {
  int nsh, nkx, nb, N;                  // A shifts (3), kx blocks stacked in N (3 or 1), rows per block, N = nkx * nb
  long long sh_stride;                  // P offset between consecutive A shifts (Wp: ky taps; (h+2) * Wp: temporal taps)
  int ktiles, nsplit, mtiles, nst;
  int cin, np, taps;                    // scratch layout dw[(tap * cin + c) * np + n], bias at dw[taps * cin * np + n]
  int ncols;                            // real output channels (n < ncols)
  int sh_center;                        // index of the unshifted A tile (1 of 3; 0 for a pointwise conv's single tile)
  int n_off;                            // first output channel of this launch inside the scratch rows (pointwise convs wider than 64)
  int arows;                            // AT rows the TMA box brings (the rows that matter: 1 + cin, rounded up to 8; <= 128).  The MMA
                                        // still reads 128 rows of the tile: the others hold stale data whose products are never stored
  int g_row0;                           // first GT row of this launch (0: the three shifted copies; kWgSpatialRows: conv5's rows)
  float* dw;
  int tmem_cols;
  int* err;
};

__device__ __forceinline__ void red_add4(float* p, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

// NC: the MMA N as a compile-time constant (instruction descriptor, accumulator columns and the lo-tile offset then cost no registers)
template <int NSH, int NC>
__global__ void __launch_bounds__(THREADS, 1) wgrad_tc_kernel(const __grid_constant__ CUtensorMap tmap_a,
                                                               const __grid_constant__ CUtensorMap tmap_g, const Params p) {
  constexpr int N = NC;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gen_base = smem_raw + (base - smem_u32(smem_raw));
  const int NST = p.nst;
  const int a_bytes = NSH * 2 * A_TILE;                   // [shift][hi|lo] tiles of 128 rows x 64 bytes
  const int b_half = N * KT * 2;                        // hi (then lo) block of N rows x 64 bytes
  const int stage_bytes = a_bytes + 2 * b_half;           // multiple of 1024: N is a multiple of 16
  const uint32_t bar_base = base + NST * stage_bytes;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (NST_MAX + s); };
  const uint32_t done_bar = bar_base + 8u * (2 * NST_MAX);
  const uint32_t tmem_slot = bar_base + 8u * (2 * NST_MAX + 1);
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(gen_base + NST * stage_bytes + 8 * (2 * NST_MAX + 1));

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  const int mt = (int)blockIdx.x % p.mtiles;
  const int split = (int)blockIdx.x / p.mtiles;
  const int per = (p.ktiles + p.nsplit - 1) / p.nsplit;
  const int k0 = split * per, k1 = min(p.ktiles, k0 + per);

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < NST_MAX; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    mbar_init(done_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, (uint32_t)p.tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  // launched through launch_chain(): barriers and tensor memory are set up while the plane builder before this kernel drains
  pdl_launch_dependents();
  pdl_wait();

  if (k0 < k1) {
    if (warp == 0) {
      if (lane == 0) {
        // ===================== TMA producer =====================
        int s = 0;
        uint32_t ph = 0;
        for (int kt = k0; kt < k1; ++kt) {
          mbar_wait(empty_bar(s), ph ^ 1u, p.err, 41);
          mbar_expect_tx(full_bar(s), (uint32_t)(NSH * 2 * p.arows * KT * 2 + 2 * b_half));
          const uint32_t st_a = base + s * stage_bytes, st_b = st_a + a_bytes;
          const int P0 = kt * KT;
          for (int sh = 0; sh < NSH; ++sh)
            for (int hl = 0; hl < 2; ++hl)
              tma_load_3d(st_a + (sh * 2 + hl) * A_TILE, &tmap_a, full_bar(s), P0 + (int)((sh - p.sh_center) * p.sh_stride), mt * ROWS, hl);
          for (int hl = 0; hl < 2; ++hl) tma_load_3d(st_b + hl * b_half, &tmap_g, full_bar(s), P0, p.g_row0, hl);
          if (++s == NST) { s = 0; ph ^= 1u; }
        }
      }
    } else if (warp == 1) {
      // ===================== MMA issuer (whole warp, one elected lane issues) =====================
      const uint32_t idesc = umma_idesc_bf16(ROWS, N);
      const uint32_t hi_sw = desc_hi(512, 4);            // K-major SWIZZLE_64B: 8-row groups 512 bytes apart
      // Everything an MMA needs beyond ONE stage-dependent descriptor base per operand is a compile-time offset (tile index, K step) or
      // hoisted out of the loop: the first version formed 18 independent descriptors per stage, the uniform registers spilled and every
      // issue paid ~4 R2UR moves (68 for 18 MMAs in the SASS; 2,300 cycles per stage against 870 of tensor time)
      const uint32_t b_lo_off = (uint32_t)b_half >> 4;
      uint32_t dcol[NSH];
#pragma unroll
      for (int sh = 0; sh < NSH; ++sh) dcol[sh] = tmem_base + (uint32_t)(sh * N);
      int s = 0;
      uint32_t ph = 0;
      for (int kt = k0; kt < k1; ++kt) {
        mbar_wait(full_bar(s), ph, p.err, 42);
        tc_fence_after();
        const uint32_t st_a = base + s * stage_bytes;
        const uint32_t a0 = desc_lo(st_a, 16), b0 = desc_lo(st_a + a_bytes, 16);
        const uint32_t first = kt > k0 ? 1u : 0u;
#pragma unroll
        for (int ks = 0; ks < KT / 16; ++ks) {
#pragma unroll
          for (int term = 0; term < 3; ++term) {
            const uint64_t bd = desc_join(b0 + (term == 1 ? b_lo_off : 0u) + 2u * ks, hi_sw);
#pragma unroll
            for (int sh = 0; sh < NSH; ++sh) {            // consecutive MMAs go to different accumulators
              const uint64_t ad = desc_join(a0 + (uint32_t)(((sh * 2 + (term == 2 ? 1 : 0)) * A_TILE) >> 4) + 2u * ks, hi_sw);
              umma_bf16_elect(dcol[sh], ad, bd, idesc, (ks > 0 || term > 0) ? 1u : first);
            }
          }
        }
        umma_commit_elect(empty_bar(s));
        if (++s == NST) { s = 0; ph ^= 1u; }
      }
      umma_commit_elect(done_bar);
    } else {
      // ===================== epilogue warps 2..5: accumulator -> fp32 scratch =====================
      const int q = warp & 3;
      const int row = mt * ROWS + q * 32 + lane;          // AT row: 0 = the ones row (bias), 1 + c = channel c
      const int c = row - 1;
      mbar_wait(done_bar, 0, p.err, 43);
      tc_fence_after();
      const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16);
      const bool is_bias = row == 0;
      const bool is_w = c >= 0 && c < p.cin;
      for (int sh = 0; sh < NSH; ++sh) {
        for (int n0 = 0; n0 < N; n0 += 16) {
          uint32_t r[16];
          tmem_ld16(lane_addr + (uint32_t)(sh * N + n0), r);      // warp-uniform
          tmem_ld_wait();
          const int kx = n0 / p.nb, nn = n0 - kx * p.nb;            // 16-column groups never straddle a kx block (nb % 16 == 0)
          const int tap = p.nkx == 3 ? sh * 3 + kx : sh;
          if (is_w) {
            float* o = p.dw + ((size_t)tap * p.cin + c) * p.np + p.n_off + nn;
#pragma unroll
            for (int j = 0; j < 16; j += 4)
              if (p.n_off + nn + j < p.np)
                red_add4(o + j, __uint_as_float(r[j]), __uint_as_float(r[j + 1]), __uint_as_float(r[j + 2]), __uint_as_float(r[j + 3]));
          } else if (is_bias && sh == p.sh_center && (p.nkx == 1 || kx == 1)) {
            float* o = p.dw + (size_t)p.taps * p.cin * p.np + p.n_off + nn;
#pragma unroll
            for (int j = 0; j < 16; j += 4)
              if (p.n_off + nn + j < p.np)
                red_add4(o + j, __uint_as_float(r[j]), __uint_as_float(r[j + 1]), __uint_as_float(r[j + 2]), __uint_as_float(r[j + 3]));
          }
        }
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

// ---- spatial convs on narrow frames: sliding window of activation tiles ----------------------------------------------------------
// The kernel above fetches three row-shifted activation tiles per stage, i.e. every tile three times (from L2: the planes fit, but at
// four septuplets per step the launches run at the L2 -> SM bandwidth, 5 TB/s, not at the tensor pipe).  When the row pitch is R <= 4
// tiles, stage s needs the tiles s - R, s, s + R of the plane: a ring of NA = 11 tiles (hi + lo, 16 KB each) holds the window, every tile
// is fetched ONCE per CTA (plus 2 R halo tiles per K range) and released after the last stage that reads it (stage index = tile index).
// The gradient tiles keep their own 3-slot ring.  Same MMAs, same accumulators, same epilogue.
constexpr int WIN_NA = 11;
constexpr int WIN_NB = 3;
constexpr int WIN_A_SLOT = 2 * A_TILE;                  // hi + lo
constexpr int WIN_B_SLOT = 2 * 96 * KT * 2;             // hi + lo blocks of 96 rows x 64 bytes
__global__ void __launch_bounds__(THREADS, 1) wgrad_tc_window_kernel(const __grid_constant__ CUtensorMap tmap_a,
                                                                      const __grid_constant__ CUtensorMap tmap_g, const Params p, const int R) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gen_base = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t b_base = base + WIN_NA * WIN_A_SLOT;
  const uint32_t bar_base = b_base + WIN_NB * WIN_B_SLOT;
  auto afull = [&](int s) { return bar_base + 8u * s; };
  auto aempty = [&](int s) { return bar_base + 8u * (WIN_NA + s); };
  auto bfull = [&](int s) { return bar_base + 8u * (2 * WIN_NA + s); };
  auto bempty = [&](int s) { return bar_base + 8u * (2 * WIN_NA + WIN_NB + s); };
  const uint32_t done_bar = bar_base + 8u * (2 * WIN_NA + 2 * WIN_NB);
  const uint32_t tmem_slot = bar_base + 8u * (2 * WIN_NA + 2 * WIN_NB + 1);
  volatile uint32_t* tmem_slot_ptr =
      reinterpret_cast<volatile uint32_t*>(gen_base + WIN_NA * WIN_A_SLOT + WIN_NB * WIN_B_SLOT + 8 * (2 * WIN_NA + 2 * WIN_NB + 1));

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  const int mt = (int)blockIdx.x % p.mtiles;
  const int split = (int)blockIdx.x / p.mtiles;
  const int per = (p.ktiles + p.nsplit - 1) / p.nsplit;
  const int k0 = split * per, k1 = min(p.ktiles, k0 + per);
  const int nt = k1 - k0;                                // stages of this CTA; A tiles i = 0 .. nt + 2R - 1 <-> plane tiles k0 - R + i

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < WIN_NA; ++s) {
      mbar_init(afull(s), 1);
      mbar_init(aempty(s), 1);
    }
    for (int s = 0; s < WIN_NB; ++s) {
      mbar_init(bfull(s), 1);
      mbar_init(bempty(s), 1);
    }
    mbar_init(done_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, (uint32_t)p.tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  const int b_half = 96 * KT * 2;

  if (nt > 0) {
    if (warp == 0) {
      if (lane == 0) {
        // ===================== TMA producer: A tile i, then the gradient tile of stage i - 2R =====================
        for (int i = 0; i < nt + 2 * R; ++i) {
          const int slot = i % WIN_NA;
          mbar_wait(aempty(slot), (uint32_t)((i / WIN_NA) & 1) ^ 1u, p.err, 44);
          mbar_expect_tx(afull(slot), (uint32_t)(2 * p.arows * KT * 2));
          const int P0 = (k0 - R + i) * KT;
          for (int hl = 0; hl < 2; ++hl) tma_load_3d(base + slot * WIN_A_SLOT + hl * A_TILE, &tmap_a, afull(slot), P0, mt * ROWS, hl);
          const int s = i - 2 * R;
          if (s >= 0) {
            const int bs = s % WIN_NB;
            mbar_wait(bempty(bs), (uint32_t)((s / WIN_NB) & 1) ^ 1u, p.err, 45);
            mbar_expect_tx(bfull(bs), (uint32_t)(2 * b_half));
            for (int hl = 0; hl < 2; ++hl) tma_load_3d(b_base + bs * WIN_B_SLOT + hl * b_half, &tmap_g, bfull(bs), (k0 + s) * KT, 0, hl);
          }
        }
      }
    } else if (warp == 1) {
      // ===================== MMA issuer =====================
      const uint32_t idesc = umma_idesc_bf16(ROWS, 96);
      const uint32_t hi_sw = desc_hi(512, 4);
      for (int s = 0; s < nt; ++s) {
        // the three tiles of this stage: i = s, s + R, s + 2R (each waited on its own barrier / phase), and its gradient tile
        uint32_t a0[3];
#pragma unroll
        for (int sh = 0; sh < 3; ++sh) {
          const int i = s + sh * R;
          mbar_wait(afull(i % WIN_NA), (uint32_t)((i / WIN_NA) & 1), p.err, 46);
          a0[sh] = desc_lo(base + (uint32_t)(i % WIN_NA) * WIN_A_SLOT, 16);
        }
        const int bs = s % WIN_NB;
        mbar_wait(bfull(bs), (uint32_t)((s / WIN_NB) & 1), p.err, 47);
        tc_fence_after();
        const uint32_t b0 = desc_lo(b_base + (uint32_t)bs * WIN_B_SLOT, 16);
        const uint32_t first = s > 0 ? 1u : 0u;
#pragma unroll
        for (int ks = 0; ks < KT / 16; ++ks) {
#pragma unroll
          for (int term = 0; term < 3; ++term) {
            const uint64_t bd = desc_join(b0 + (term == 1 ? (uint32_t)(96 * KT * 2 >> 4) : 0u) + 2u * ks, hi_sw);
#pragma unroll
            for (int sh = 0; sh < 3; ++sh) {
              const uint64_t ad = desc_join(a0[sh] + (term == 2 ? (uint32_t)(A_TILE >> 4) : 0u) + 2u * ks, hi_sw);
              umma_bf16_elect(tmem_base + (uint32_t)(sh * 96), ad, bd, idesc, (ks > 0 || term > 0) ? 1u : first);
            }
          }
        }
        umma_commit_elect(aempty(s % WIN_NA));          // tile i = s: this was its last reader
        umma_commit_elect(bempty(bs));
      }
      umma_commit_elect(done_bar);
    } else {
      // ===================== epilogue warps 2..5 =====================
      const int q = warp & 3;
      const int row = mt * ROWS + q * 32 + lane;
      const int c = row - 1;
      mbar_wait(done_bar, 0, p.err, 48);
      tc_fence_after();
      const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16);
      const bool is_bias = row == 0;
      const bool is_w = c >= 0 && c < p.cin;
      for (int sh = 0; sh < 3; ++sh) {
        for (int n0 = 0; n0 < 96; n0 += 16) {
          uint32_t r[16];
          tmem_ld16(lane_addr + (uint32_t)(sh * 96 + n0), r);
          tmem_ld_wait();
          const int kx = n0 / 32, nn = n0 - kx * 32;
          const int tap = sh * 3 + kx;
          if (is_w) {
            float* o = p.dw + ((size_t)tap * p.cin + c) * p.np + nn;
#pragma unroll
            for (int j = 0; j < 16; j += 4)
              red_add4(o + j, __uint_as_float(r[j]), __uint_as_float(r[j + 1]), __uint_as_float(r[j + 2]), __uint_as_float(r[j + 3]));
          } else if (is_bias && sh == 1 && kx == 1) {
            float* o = p.dw + (size_t)p.taps * p.cin * p.np + nn;
#pragma unroll
            for (int j = 0; j < 16; j += 4)
              red_add4(o + j, __uint_as_float(r[j]), __uint_as_float(r[j + 1]), __uint_as_float(r[j + 2]), __uint_as_float(r[j + 3]));
          }
        }
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

// ---- plane builders ------------------------------------------------------------------------------------------------
__device__ __forceinline__ long long plane_index(long long m, const WgGeom& g) {
  const long long hw = (long long)g.h * g.w;
  const long long n = m / hw, pix = m - n * hw;
  const int y = (int)(pix / g.w), x = (int)(pix - (long long)y * g.w);
  const long long b = n / g.T, t = n - b * g.T;
  const long long f = b * (g.T + 1) + t + 1;
  return (f * (g.h + 2) + y + 1) * g.Wp + x + 1;
}

// activations: slab-planar (hi, lo) dense buffer [pitch/16][M][16 hi | 16 lo] -> AT planes [2][193][Pa]; grid (pixels / 256, slabs)
__global__ void __launch_bounds__(256) wg_planes_act_kernel(const __nv_bfloat16* __restrict__ buf, long long M, __nv_bfloat16* __restrict__ at,
                                                            const WgGeom g) {
  chain_entry();
  const long long m = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= M) return;
  const int slab = blockIdx.y;
  const long long P = plane_index(m, g);
  const uint4* src = reinterpret_cast<const uint4*>(buf + ((size_t)slab * M + m) * 32);
  uint4 r[4] = {__ldg(src), __ldg(src + 1), __ldg(src + 2), __ldg(src + 3)};      // 16 hi, 16 lo
  const uint16_t* e = reinterpret_cast<const uint16_t*>(r);
  uint16_t* hi = reinterpret_cast<uint16_t*>(at) + (size_t)(1 + slab * 16) * g.Pa + P;
  uint16_t* lo = hi + (size_t)kWgRows * g.Pa;
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    hi[(size_t)j * g.Pa] = e[j];
    lo[(size_t)j * g.Pa] = e[16 + j];
  }
  if (slab == 0) {
    reinterpret_cast<uint16_t*>(at)[P] = 0x3F80;      // row 0 of the hi plane: 1.0 at real pixels (lo stays 0)
  }
}

// activations of a pointwise conv: fp32 pixel-major [M][pitch], channels [0, C) -> AT rows 1 + c (one thread per (4 channels, pixel))
__global__ void __launch_bounds__(256) wg_planes_act_f32_kernel(const float* __restrict__ src, int pitch, int C, long long M,
                                                                __nv_bfloat16* __restrict__ at, const WgGeom g) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= M * (C / 4)) return;
  const int c0 = (int)(idx / M) * 4;
  const long long m = idx - (long long)(c0 / 4) * M;
  const long long P = plane_index(m, g);
  const float4 t = __ldg(reinterpret_cast<const float4*>(src + m * pitch + c0));
  const float v4[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    __nv_bfloat16 h, l;
    x2_split(v4[e], h, l);
    __nv_bfloat16* hi = at + (size_t)(1 + c0 + e) * g.Pa + P;
    hi[0] = h;
    hi[(size_t)kWgRows * g.Pa] = l;
  }
  if (c0 == 0) reinterpret_cast<uint16_t*>(at)[P] = 0x3F80;      // the ones row
}

// gradient: fp32 pixel-major g[m * pitch + off + n], n < ncols -> GT planes [2][160][Pa].  Spatial convs (ncopies = 3): rows
// kx * 32 + n hold the gradient written at P + (kx - 1), i.e. GT_kx[n][P] = g[n][P - (kx-1)] (the shifted positions are padding
// columns of the same row, so the three copies never collide with real pixels of a neighbouring row); conv5 (ncopies = 1): rows
// 96 + n at P, rows >= ncols zero.  The two forms use disjoint rows: each row is always written at the same set of positions.
__global__ void __launch_bounds__(256) wg_planes_grad_kernel(const float* __restrict__ gsrc, int pitch, int off, long long sslabM, int ncols, int nb,
                                                             int ncopies, long long M, __nv_bfloat16* __restrict__ gt, const WgGeom g,
                                                             float* __restrict__ zero, long long zero_n) {
  // one thread per (4-channel group, pixel), pixels fastest: a warp writes 64 contiguous bytes per plane row and store instruction,
  // and the launch has nb / 4 times the threads of a thread-per-pixel form (at 50 K pixels that form left the SMs a third full)
  chain_entry();
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  // the split-K scratch of the weight-gradient launch that follows (one memset node per conv less in the chain)
  for (long long i = idx; i < zero_n; i += (long long)gridDim.x * blockDim.x) zero[i] = 0.f;
  if (idx >= M * (nb / 4)) return;
  const int n0 = (int)(idx / M) * 4;
  const long long m = idx - (long long)(n0 / 4) * M;
  const long long P = plane_index(m, g);
  const int row0 = ncopies == 3 ? 0 : kWgSpatialRows;
  const float* src = gsrc + dense_off(m, off + n0, pitch, sslabM);      // 4-channel groups never straddle a 16-channel slab
  float v4[4];
  if (((pitch | off) & 3) == 0 && n0 + 4 <= ncols) {
    const float4 t = __ldg(reinterpret_cast<const float4*>(src));
    v4[0] = t.x; v4[1] = t.y; v4[2] = t.z; v4[3] = t.w;
  } else {
#pragma unroll
    for (int e = 0; e < 4; ++e) v4[e] = n0 + e < ncols ? __ldg(src + e) : 0.f;
  }
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    __nv_bfloat16 h, l;
    x2_split(v4[e], h, l);
    for (int kx = 0; kx < ncopies; ++kx) {
      __nv_bfloat16* hi = gt + (size_t)(row0 + kx * nb + n0 + e) * g.Pa + P + (ncopies == 3 ? kx - 1 : 0);
      hi[0] = h;
      hi[(size_t)kWgGradRows * g.Pa] = l;
    }
  }
}

}  // namespace wg

bool train_pdl_enabled() {
  static int on = -1;
  if (on < 0) {
    const char* e = getenv("SELFC_TRAIN_PDL");
    on = (e && atoi(e) == 0) ? 0 : 1;
  }
  return on == 1;
}

// SELFC_WGRAD_WINDOW: the sliding-window form of the spatial kernel (and the 32-pixel row pitch it needs)
static bool wg_window_enabled() {
  static int on = -1;
  if (on < 0) {
    const char* e = getenv("SELFC_WGRAD_WINDOW");
    on = e ? (atoi(e) != 0 ? 1 : 0) : kWgWindowDefault;
  }
  return on == 1;
}

WgGeom wg_geometry(const Dims& d) {
  WgGeom g;
  g.B = d.B; g.T = d.T; g.h = d.h; g.w = d.w;
  // row pitch: a multiple of 8 pixels (16-byte TMA coordinates); narrow frames take a multiple of 32 so that a row shift is a whole
  // number R <= 4 of 32-pixel tiles and the spatial kernel can keep a sliding window of activation tiles (wgrad_tc_window_kernel)
  const int wp32 = (d.w + 2 + 31) & ~31;
  g.Wp = wg_window_enabled() && wp32 / 32 <= kWgWindowMaxR ? wp32 : ((d.w + 2 + 7) & ~7);
  g.Fp = (long long)(d.h + 2) * g.Wp;
  g.P = ((long long)d.B * (d.T + 1) + 1) * g.Fp;
  g.Pa = (g.P + 31) & ~31ll;
  return g;
}

size_t wg_plane_bytes(const WgGeom& g) { return (size_t)2 * (kWgRows + kWgGradRows) * g.Pa * sizeof(__nv_bfloat16); }

int launch_wg_planes_act(const bfx2* buf, int pitch, const Dims& d, const WgGeom& g, void* planes, cudaStream_t st) {
  SELFC_CHECK_ARG(pitch % 16 == 0 && pitch <= kWgRows - 1, "wgrad planes: pitch %d", pitch);
  const long long M = d.M();
  dim3 grid(cdiv(M, 256), pitch / 16);
  SELFC_CUDA(launch_chain(wg::wg_planes_act_kernel, grid, 256, 0, st, reinterpret_cast<const __nv_bfloat16*>(buf), M,
                          reinterpret_cast<__nv_bfloat16*>(planes), g));
  SELFC_LAUNCH_CHECK("wg_planes_act_kernel");
  return 0;
}

int launch_wg_planes_act_f32(const float* src, int pitch, int C, const Dims& d, const WgGeom& g, void* planes, cudaStream_t st) {
  SELFC_CHECK_ARG(C % 4 == 0 && pitch % 4 == 0 && C <= kWgRows - 1 && aligned16(src), "wgrad planes: %d fp32 channels (pitch %d)", C, pitch);
  wg::wg_planes_act_f32_kernel<<<cdiv(d.M() * (C / 4), 256), 256, 0, st>>>(src, pitch, C, d.M(), reinterpret_cast<__nv_bfloat16*>(planes), g);
  SELFC_LAUNCH_CHECK("wg_planes_act_f32_kernel");
  return 0;
}

int launch_wg_planes_grad(const float* gsrc, int pitch, int off, long long sslabM, int ncols, int nb, bool temporal, const Dims& d,
                          const WgGeom& g, void* planes, cudaStream_t st, float* zero, long long zero_n) {
  SELFC_CHECK_ARG(nb % 16 == 0 && nb <= kWgGradRows - kWgSpatialRows && ncols <= nb && (temporal || nb == 32),
                  "wgrad planes: %d gradient columns", ncols);
  __nv_bfloat16* gt = reinterpret_cast<__nv_bfloat16*>(planes) + (size_t)2 * kWgRows * g.Pa;
  SELFC_CUDA(launch_chain(wg::wg_planes_grad_kernel, dim3((unsigned)cdiv(d.M() * (nb / 4), 256)), 256, 0, st, gsrc, pitch, off, sslabM, ncols, nb,
                          temporal ? 1 : 3, d.M(), gt, g, zero, zero_n));
  SELFC_LAUNCH_CHECK("wg_planes_grad_kernel");
  return 0;
}

int launch_wgrad_tc(void* planes, const WgGeom& g, int cin, int ncols, int nb, int kind, float* dw, int np, cudaStream_t st, int n_off) {
  const bool temporal = kind != WG_SPATIAL;        // conv5 and the pointwise convs read the unshifted gradient rows
  const int taps = kind == WG_SPATIAL ? 9 : (kind == WG_TEMPORAL ? 3 : 1);
  SELFC_CHECK_ARG(kind >= WG_SPATIAL && kind <= WG_POINT && nb % 16 == 0 && nb >= 16 && nb <= kWgGradRows - kWgSpatialRows && cin >= 1 &&
                      cin <= kWgRows - 1 && np % 4 == 0 && n_off % 4 == 0 && n_off >= 0 && (temporal || nb == 32),
                  "wgrad_tc: unsupported shape (cin %d, nb %d, kind %d)", cin, nb, kind);
  tc::EncodeTiledFn encode = tc::get_encode_fn();
  if (!encode) {
    set_error("cuTensorMapEncodeTiled is not available from this driver");
    return SELFC_E_CUDA;
  }
  __nv_bfloat16* at = reinterpret_cast<__nv_bfloat16*>(planes);
  __nv_bfloat16* gt = at + (size_t)2 * kWgRows * g.Pa;
  CUtensorMap tmap_a, tmap_g;
  const cuuint32_t estr[3] = {1, 1, 1};
  // one M tile: only the rows that exist are fetched (cin = 16 -> 24 of 128); two M tiles: full boxes (the second is zero-filled past row 192)
  const int arows = cin + 1 <= wg::ROWS ? ((cin + 1 + 7) & ~7) : wg::ROWS;
  {
    const cuuint64_t gdim[3] = {(cuuint64_t)g.P, (cuuint64_t)kWgRows, 2};
    const cuuint64_t gstr[2] = {(cuuint64_t)g.Pa * 2, (cuuint64_t)kWgRows * g.Pa * 2};
    const cuuint32_t box[3] = {(cuuint32_t)wg::KT, (cuuint32_t)arows, 1};
    CUresult r = encode(&tmap_a, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, at, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      set_error("cuTensorMapEncodeTiled (wgrad activations) failed with CUresult %d", (int)r);
      return SELFC_E_CUDA;
    }
  }
  {
    const cuuint64_t gdim[3] = {(cuuint64_t)g.P, (cuuint64_t)kWgGradRows, 2};
    const cuuint64_t gstr[2] = {(cuuint64_t)g.Pa * 2, (cuuint64_t)kWgGradRows * g.Pa * 2};
    const cuuint32_t box[3] = {(cuuint32_t)wg::KT, (cuuint32_t)(temporal ? nb : 3 * nb), 1};
    CUresult r = encode(&tmap_g, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, gt, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      set_error("cuTensorMapEncodeTiled (wgrad gradient) failed with CUresult %d", (int)r);
      return SELFC_E_CUDA;
    }
  }
  wg::Params p;
  memset(&p, 0, sizeof(p));
  p.nsh = kind == WG_POINT ? 1 : 3;
  p.sh_center = kind == WG_POINT ? 0 : 1;
  p.n_off = n_off;
  p.nkx = temporal ? 1 : 3;
  p.nb = nb;
  p.N = p.nkx * nb;
  p.sh_stride = kind == WG_TEMPORAL ? g.Fp : (kind == WG_SPATIAL ? g.Wp : 0);
  p.g_row0 = temporal ? kWgSpatialRows : 0;
  p.arows = arows;
  p.ktiles = (int)(g.Pa / wg::KT);
  p.mtiles = cdiv(cin + 1, wg::ROWS);
  const int nsm = tc::num_sms();
  // one wave of CTAs: the per-CTA cost that does not shrink with the split is the reduction of its 128 x 288 accumulator into the
  // scratch (SELFC_WGRAD_WAVES=2 doubles the split)
  static int waves = 0;
  if (!waves) {
    const char* e = getenv("SELFC_WGRAD_WAVES");
    waves = e ? atoi(e) : 1;
    if (waves < 1 || waves > 8) waves = 1;
  }
  int nsplit = waves * nsm / p.mtiles;
  if (nsplit > p.ktiles / 4) nsplit = p.ktiles / 4;      // at least four K tiles per CTA
  if (nsplit < 1) nsplit = 1;
  p.nsplit = nsplit;
  p.cin = cin; p.np = np; p.taps = taps; p.ncols = ncols; p.dw = dw;
  int cols = p.nsh * p.N, pw = 32;
  while (pw < cols) pw <<= 1;
  p.tmem_cols = pw;
  p.err = tc::err_flag_for_device();
  const int stage_bytes = p.nsh * 2 * wg::A_TILE + 2 * p.N * wg::KT * 2;
  int nst = (227 * 1024 - wg::BAR_BYTES - 1024) / stage_bytes;
  if (nst > wg::NST_MAX) nst = wg::NST_MAX;
  p.nst = nst;
  const int smem = nst * stage_bytes + wg::BAR_BYTES + 1024;
  static bool smem_set[64] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev >= 0 && dev < 64 && !smem_set[dev]) {
    SELFC_CUDA(cudaFuncSetAttribute(wg::wgrad_tc_kernel<3, 96>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    SELFC_CUDA(cudaFuncSetAttribute(wg::wgrad_tc_kernel<3, 64>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    SELFC_CUDA(cudaFuncSetAttribute(wg::wgrad_tc_kernel<3, 48>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    SELFC_CUDA(cudaFuncSetAttribute(wg::wgrad_tc_kernel<3, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    SELFC_CUDA(cudaFuncSetAttribute(wg::wgrad_tc_kernel<1, 64>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    SELFC_CUDA(cudaFuncSetAttribute(wg::wgrad_tc_kernel<1, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    smem_set[dev] = true;
  }
  // spatial convs on narrow frames: the sliding-window form (SELFC_WGRAD_WINDOW=0 / 1: the three-fetch form / the window form)
  if (wg_window_enabled() && kind == WG_SPATIAL && g.Wp % 32 == 0 && g.Wp / 32 <= kWgWindowMaxR && np == 32) {
    const int wsmem = wg::WIN_NA * wg::WIN_A_SLOT + wg::WIN_NB * wg::WIN_B_SLOT + wg::BAR_BYTES + 1024;
    static bool wset[64] = {};
    if (dev >= 0 && dev < 64 && !wset[dev]) {
      SELFC_CUDA(cudaFuncSetAttribute(wg::wgrad_tc_window_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
      wset[dev] = true;
    }
    wg::wgrad_tc_window_kernel<<<p.mtiles * p.nsplit, wg::THREADS, wsmem, st>>>(tmap_a, tmap_g, p, g.Wp / 32);
    SELFC_LAUNCH_CHECK("wgrad_tc_window_kernel");
    return 0;
  }
  const int grid = p.mtiles * p.nsplit;
  if (p.nsh == 3 && p.N == 96) SELFC_CUDA(launch_chain(wg::wgrad_tc_kernel<3, 96>, dim3((unsigned)grid), wg::THREADS, smem, st, tmap_a, tmap_g, p));
  else if (p.nsh == 3 && p.N == 64) SELFC_CUDA(launch_chain(wg::wgrad_tc_kernel<3, 64>, dim3((unsigned)grid), wg::THREADS, smem, st, tmap_a, tmap_g, p));
  else if (p.nsh == 3 && p.N == 48) SELFC_CUDA(launch_chain(wg::wgrad_tc_kernel<3, 48>, dim3((unsigned)grid), wg::THREADS, smem, st, tmap_a, tmap_g, p));
  else if (p.nsh == 3 && p.N == 16) SELFC_CUDA(launch_chain(wg::wgrad_tc_kernel<3, 16>, dim3((unsigned)grid), wg::THREADS, smem, st, tmap_a, tmap_g, p));
  else if (p.nsh == 1 && p.N == 64) SELFC_CUDA(launch_chain(wg::wgrad_tc_kernel<1, 64>, dim3((unsigned)grid), wg::THREADS, smem, st, tmap_a, tmap_g, p));
  else if (p.nsh == 1 && p.N == 16) SELFC_CUDA(launch_chain(wg::wgrad_tc_kernel<1, 16>, dim3((unsigned)grid), wg::THREADS, smem, st, tmap_a, tmap_g, p));
  else {
    set_error("wgrad_tc: no instantiation for %d shifts x N = %d", p.nsh, p.N);
    return SELFC_E_UNSUPPORTED;
  }
  SELFC_LAUNCH_CHECK("wgrad_tc_kernel");
  return 0;
}

}  // namespace selfc
