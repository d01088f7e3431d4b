// (3,1,1) temporal convolution of D2DTInput (conv5) with the InvBlockExp coupling fused into its epilogue, and the
// GlobalAgg apply step, as one tcgen05 / TMEM / TMA kernel (BF16 mode).
//
// Reference behaviour restated (not copied):
//   conv5                      Subnet_constructor.py:106,130     out[t] = sum_{dt} W[dt] . in[t+dt-1], zero-padded at clip ends
//   coupling                   SelfC_GMM_arch_inv.py:21-33      y1 = x1 +/- F, s = 2*sigmoid(H)-1, y2 = x2*e^s + G | (x2-G)/e^s
//   GlobalAgg apply            SelfC_GMM_arch_inv.py:266,278-285 out[t'] = x[t'] + sum_t W[b,t,t'] * proj1(x[t])
//
// Mapping: a CTA owns 128 consecutive pixels of ALL T frames of one clip and walks the INPUT frames in order.  The tile
// of input frame f (one ring stage per K chunk, loaded once, released as soon as its MMAs are committed) is the A operand
// of up to three MMAs per K step issued back to back -- out[f-1] (tap 2), out[f] (tap 1), out[f+1] (tap 0).  Output
// frames live in ROLLING TMEM accumulators (4 x N columns): out[f-1] is complete when frame f has been consumed, so its
// epilogue runs while the MMAs of the following frames are issued, across tile boundaries too, and T is not limited by
// TMEM.  Clip-end zero padding = the corresponding MMAs are simply not issued.  Weights stay resident in shared memory.
// Input: a slab-planar dense buffer (common.cuh; one 4-D TMA box of 128 px x kps slabs per stage, SWIZZLE_32B sub-tiles
// of 4 KB) or a pixel-major buffer (GMM head; 3-D box of 128 px x 64 ch, SWIZZLE_128B rows).
// Epilogue warps read the accumulator of the finished pass (tcgen05.ld), apply bias and the fused coupling arithmetic on
// the fp32 latent state, and release the accumulator.
#include <cuda.h>
#include <cuda_fp16.h>
#include <stdlib.h>

#include "common.cuh"
#include "conv_tc.h"
#include "kernels.h"
#include "pack_batch.h"
#include "tc_ptx.cuh"

namespace selfc {
namespace tc5 {

using namespace tc;

constexpr int MT = 128;
constexpr int NST_MAX = 14;      // ring slots
constexpr int NACC_MAX = 4;      // rolling accumulators
constexpr int THREADS = 192;
constexpr int BAR_BYTES = 512;
constexpr int BIAS_BYTES = 1024;   // up to 256 fp32

struct Params {
  const void* wimg;
  const float* bias;
  const void* wimg2;                          // EPI_COUPLE_HG: second conv (G); its chunks follow the first conv's in a frame
  const float* bias2;
  int nc_split;                               // chunks per frame that belong to the first conv / tensor map (== nchunk: single conv)
  int nhalf;                                  // columns of one conv inside the accumulator (npad: single conv; npad/2: HG)
  int T, B, hw, nks, npad, taps, cout, nst;   // nks = K steps of 16 channels; nst = ring slots
  int kps, nchunk, nacc;                      // K steps per stage, stages per frame, rolling accumulators
  int epi_quads;                              // 16-byte quads of epilogue operands staged per pixel-frame (Y2: 24, Y1: 1)
  int tiles_p, ntiles, tmem_cols;
  int zrev;         // walk the pixel strips last-to-first (tc::next_direction)
  long long m_limit;   // rows (pixels) that really exist in the buffer
  int epi, rev, act;
  int in_slab;                                // input is a slab-planar dense buffer (common.cuh): 4-D TMA, SWIZZLE_32B sub-tiles
  __nv_bfloat16* outT;
  int outT_pitch, outT_off;
  long long outT_slabM, copy_slabM;           // layouts of outT and of copyA / copyB (0 = pixel-major)
  float* outF;
  int outF_pitch, outF_off, outF_planar;   // outF_planar: write quads [C/4][m_limit][4] instead of [M][pitch]; 2: the quads are fp16
  int outF_blk, outF_blk_stride;           // planar: column n -> channel outF_off + (n / blk) * stride + n % blk (blk = 0: outF_off + n)
  long long outF_slabM;                    // non-planar outF as 16-channel fp32 slabs (0: pixel-major)
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
  int gmm_k, gmm_T;     // EPI_GMM: mixture component of this launch, frames per clip (eps index)
  long long gmm_hw;     //          pixels per frame
  const float* eps;     //          injected noise [B,48,5,T,h,w] or null
  uint64_t seed, offset;
  int* err;
  long long* dbg;   // optional: CTA 0 writes cycles waited per barrier site [16] + totals
};

// Stage the coupling operands of the frame `ahead` frames after (tile, t) -- in this CTA's processing order -- into the
// calling thread's own shared-memory slots with cp.async, and commit one group (always, so group counting is uniform).
// Y2: quads 0..11 = x2 (latent state quads 1..12), 12..23 = log-scale s.  Y1: quad 0 = x1.
__device__ __forceinline__ void stage_epilogue_operands(const Params& p, int tile, int t, int ahead, int row, size_t Mtot, uint32_t slot) {
  int t2 = t + ahead, tile2 = tile;
  while (t2 >= p.T) { t2 -= p.T; tile2 += (int)gridDim.x; }
  if (tile2 < p.ntiles) {
    if (p.zrev) tile2 = p.ntiles - 1 - tile2;
    const int b = tile2 / p.tiles_p;
    const int pix = (tile2 - b * p.tiles_p) * MT + row;
    if (pix < p.hw) {
      const size_t m = ((size_t)b * p.T + t2) * p.hw + pix;
      if (p.epi == EPI_COUPLE_HG) {
#pragma unroll
        for (int qd = 0; qd < kSQuads; ++qd)
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(slot + (uint32_t)(qd * MT) * 16u), "l"(p.z + quad_off(Mtot, 1 + qd, m)) : "memory");
      } else if (p.epi == EPI_COUPLE_Y2) {
#pragma unroll
        for (int qd = 0; qd < kSQuads; ++qd) {
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(slot + (uint32_t)(qd * MT) * 16u), "l"(p.z + quad_off(Mtot, 1 + qd, m)) : "memory");
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(slot + (uint32_t)((kSQuads + qd) * MT) * 16u), "l"(p.sbuf + quad_off(Mtot, qd, m)) : "memory");
        }
      } else {
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(slot), "l"(p.z + quad_off(Mtot, 0, m)) : "memory");
      }
    }
  }
  asm volatile("cp.async.commit_group;" ::: "memory");
}

// barrier wait; with -DSELFC_TC_TIMING the cycles spent waiting are accumulated for the debug counters
__device__ __forceinline__ void timed_wait(uint32_t bar, uint32_t parity, int* err, int code, long long& acc) {
#ifdef SELFC_TC_TIMING
  const long long t0 = clock64();
  mbar_wait(bar, parity, err, code);
  acc += clock64() - t0;
#else
  mbar_wait(bar, parity, err, code);
#endif
}

__device__ __forceinline__ uint32_t pack_bf2(float a, float b) {
  __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ void store_bf16x8(__nv_bfloat16* o, const float* v) {
  uint4 pk;
  pk.x = pack_bf2(v[0], v[1]); pk.y = pack_bf2(v[2], v[3]); pk.z = pack_bf2(v[4], v[5]); pk.w = pack_bf2(v[6], v[7]);
  *reinterpret_cast<uint4*>(o) = pk;
}
__device__ __forceinline__ void unpack_bf16x8(const uint4 r, float* v) {
  const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    __nv_bfloat162 b = *reinterpret_cast<const __nv_bfloat162*>(&w[i]);
    v[2 * i] = __low2float(b);
    v[2 * i + 1] = __high2float(b);
  }
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
// streaming 16-byte load that the compiler may hoist above later stores (the buffers never alias within a launch)
__device__ __forceinline__ float4 ld_nc4(const float* p) {
  float4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
  return v;
}
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t r[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr)
               : "memory");
}

__device__ __forceinline__ float4 lds4(uint32_t saddr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(saddr));
  return v;
}
__device__ __forceinline__ void cp_async16(uint32_t saddr, const void* g) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(saddr), "l"(g) : "memory");
}

// X2 (BF16X3 mode, common.cuh): activations and weights are (hi, lo) bf16 pairs.  16 channels of a pixel are one 64-byte row
// [16 x hi | 16 x lo] -- a SWIZZLE_64B sub-tile row of a slab, or a quarter of a SWIZZLE_128B row of a pixel-major buffer --
// every (tap, K-step) has a hi and a lo weight tile, and a K-step issues three MMAs per tap (hi.hi, hi.lo, lo.hi).
template <int TAPS, bool GMM, bool X2 = false>
__global__ void __launch_bounds__(THREADS, 1) temporal_tc_kernel(const __grid_constant__ CUtensorMap tmap,
                                                                  const __grid_constant__ CUtensorMap tmap2, const Params p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gen_base = smem_raw + (base - smem_u32(smem_raw));
  const int T = p.T;
  const int NST = p.nst;
  const int KPS = p.kps;
  const int NC = p.nchunk;
  const int NACC = p.nacc;
  const int STAGE_BYTES = KPS * MT * (X2 ? 64 : 32);   // KPS K-steps of [128 px][16 ch]
  const uint32_t a_base = base;
  const uint32_t bar_base = base + NST * STAGE_BYTES;
  const uint32_t bias_off = NST * STAGE_BYTES + BAR_BYTES;
  const uint32_t w_base = base + bias_off + BIAS_BYTES;
  // epilogue operand staging (double buffered, each thread owns its 16-byte slots): [2][epi_quads][128 px][16 B]
  const uint32_t epi_base = w_base + (uint32_t)p.taps * p.nks * p.nhalf * 32u * (p.nc_split < p.nchunk ? 2u : 1u) * (X2 ? 2u : 1u);
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (NST_MAX + s); };
  const uint32_t w_bar = bar_base + 8u * (2 * NST_MAX);
  auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * NST_MAX + 1 + a); };
  auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * NST_MAX + 1 + NACC_MAX + a); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * NST_MAX + 1 + 2 * NACC_MAX);
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(gen_base + NST * STAGE_BYTES + 8 * (2 * NST_MAX + 1 + 2 * NACC_MAX));
  float* sbias = reinterpret_cast<float*>(gen_base + bias_off);

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);   // warp-uniform for the compiler
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < NST_MAX; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    mbar_init(w_bar, 1);
    for (int a = 0; a < NACC_MAX; ++a) {
      mbar_init(tfull_bar(a), 1);
      mbar_init(tempty_bar(a), 4);
    }
    fence_barrier_init();
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap) : "memory");
  }
  if (warp == 1) tmem_alloc(tmem_slot, (uint32_t)p.tmem_cols);
  for (int i = threadIdx.x; i < p.npad; i += THREADS) sbias[i] = i < p.nhalf ? __ldg(p.bias + i) : __ldg(p.bias2 + (i - p.nhalf));
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  pdl_launch_dependents();

  const int nks = p.nks, npad = p.npad;
  constexpr int taps = TAPS;
  const int nconv = p.nhalf;                          // N of one MMA (== npad unless two convs share the accumulator)
  const uint32_t wtile = (uint32_t)nconv * 32u;
  const uint32_t wkstep = X2 ? 2u * wtile : wtile;   // bytes of weights per K-step (X2: hi tile, then lo tile)

  if (warp == 0) {
    if (lane == 0) {
      // ===================== TMA producer: stages in (tile, frame, chunk) order =====================
      const uint32_t wbytes = (uint32_t)taps * nks * wkstep;
      const bool two = p.nc_split < NC;
      mbar_expect_tx(w_bar, two ? 2u * wbytes : wbytes);
      for (int tap = 0; tap < taps; ++tap) {
        bulk_g2s(w_base + tap * nks * wkstep, (const uint8_t*)p.wimg + (size_t)tap * nks * wkstep, (uint32_t)nks * wkstep, w_bar);
        if (two)
          bulk_g2s(w_base + wbytes + tap * nks * wkstep, (const uint8_t*)p.wimg2 + (size_t)tap * nks * wkstep, (uint32_t)nks * wkstep, w_bar);
      }
      pdl_wait();      // weights are static; activations come from the previous kernel in the stream
      int s = 0;
      uint32_t ph = 0;
      long long w_prod = 0;
      const long long t_start = clock64();
      for (int tile_i = blockIdx.x; tile_i < p.ntiles; tile_i += gridDim.x) {
        const int tile = p.zrev ? p.ntiles - 1 - tile_i : tile_i;
        const int b = tile / p.tiles_p;
        const int p0 = (tile - b * p.tiles_p) * MT;
        for (int f = 0; f < T; ++f) {
          for (int c = 0; c < NC; ++c) {
            timed_wait(empty_bar(s), ph ^ 1u, p.err, 11, w_prod);
            mbar_expect_tx(full_bar(s), (uint32_t)STAGE_BYTES);
            if (c >= p.nc_split) tma_load_4d(a_base + s * STAGE_BYTES, &tmap2, full_bar(s), 0, p0, b * T + f, (c - p.nc_split) * KPS);
            else if (p.in_slab) tma_load_4d(a_base + s * STAGE_BYTES, &tmap, full_bar(s), 0, p0, b * T + f, c * KPS);
            else tma_load_3d(a_base + s * STAGE_BYTES, &tmap, full_bar(s), c * KPS * (X2 ? 32 : 16), p0, b * T + f);
            if (++s == NST) { s = 0; ph ^= 1u; }
          }
        }
      }
      if (p.dbg && blockIdx.x == 0) { p.dbg[0] = w_prod; p.dbg[1] = clock64() - t_start; }
    }
  } else if (warp == 1) {
    {
      // ===================== MMA issuer: whole warp runs the loop, one elected lane issues =====================
      const uint32_t idesc = umma_idesc_bf16(128, nconv);
      // pixel-major input: one SWIZZLE_128B row of 64 channels per pixel, a K step advances 32 bytes inside the row;
      // slab input: KPS sub-tiles of [128 px][16 ch] (SWIZZLE_32B, 4 KB each), a K step advances one sub-tile
      // X2: SWIZZLE_64B sub-tiles of 8 KB / 64 bytes per K-step inside the SWIZZLE_128B row; the lo operand is 32 bytes after the hi one
      const uint32_t hi_a = p.in_slab ? (X2 ? desc_hi(512, 4) : desc_hi(256, 6)) : desc_hi(1024, 2);
      const uint32_t a_kinc = p.in_slab ? (uint32_t)(MT * (X2 ? 64 : 32) >> 4) : (X2 ? 4u : 2u);
      long long w_full = 0, w_tempty = 0;
      const long long t_start = clock64();
      mbar_wait(w_bar, 0, p.err, 12);
      const long long t_w = clock64() - t_start;
      int gbase = 0;            // output frames of earlier tiles: output o of this tile uses accumulator (gbase + o) & (NACC-1)
      int s = 0;                // stage ring position (plain FIFO: every stage is consumed once)
      uint32_t ph = 0;
      const uint32_t amask = (uint32_t)(NACC - 1);          // NACC is 2 or 4
      const uint32_t hi_b = desc_hi(128, 0);
      const uint32_t b_lbo = ((uint32_t)nconv * 16u >> 4) << 16;
      const uint32_t wconv = (uint32_t)taps * nks * wkstep;   // bytes of one conv's weight image
      const uint32_t wstep = wkstep >> 4;                   // descriptor units per 16-channel K-step of weights
      for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x, gbase += T) {
        // INPUT-frame-major: the tile of input frame fi feeds the outputs fi-1, fi, fi+1 (taps 2, 1, 0) back to back with
        // the same A descriptor, then is released.  Output fi-1 is complete once frame fi has been consumed.
        for (int fi = 0; fi < T; ++fi) {
          // per tap: output frame, accumulator column, whether the accumulator already holds a partial sum.  Everything
          // the K loop needs is computed here so that the issue loop is a handful of uniform-datapath instructions per MMA
          // (a runtime modulo inside it cost ~100 cycles per MMA, 3x the tensor-core time of these small-N MMAs).
          uint32_t dcol[TAPS], bdesc0[TAPS], started[TAPS];
          bool valid[TAPS];
#pragma unroll
          for (int j = 0; j < TAPS; ++j) {
            const int tap = TAPS - 1 - j;                                  // issue order: taps 2, 1, 0 -> outputs fi-1, fi, fi+1
            const int o = TAPS == 3 ? fi - tap + 1 : fi;                   // out[o] += W[tap] . in[o + tap - 1]
            valid[j] = o >= 0 && o < T;
            dcol[j] = tmem_base + (((uint32_t)(gbase + o)) & amask) * (uint32_t)npad;
            bdesc0[j] = (((w_base + (uint32_t)(tap * nks) * wkstep) & 0x3FFFFu) >> 4) | b_lbo;
            // first contribution to out[o] comes from frame max(o-1, 0) (frame o for a pointwise conv)
            started[j] = TAPS == 3 ? (fi != (o > 0 ? o - 1 : 0) ? 1u : 0u) : 0u;
            if (valid[j] && !started[j]) {
              const int g = gbase + o;
              timed_wait(tempty_bar((int)((uint32_t)g & amask)), (((uint32_t)g / (uint32_t)NACC) & 1u) ^ 1u, p.err, 13, w_tempty);
            }
          }
          tc_fence_after();
          uint32_t kk = 0;                                                // K step of the current conv within the frame
          uint32_t cvoff_d = 0, cvoff_b = 0;                              // second conv: accumulator column / weight image offsets
          for (int c = 0; c < NC; ++c) {
            if (c == p.nc_split) { kk = 0; cvoff_d = (uint32_t)nconv; cvoff_b = wconv >> 4; }
            const int lc = c >= p.nc_split ? c - p.nc_split : c;
            timed_wait(full_bar(s), ph, p.err, 14, w_full);
            tc_fence_after();
            const int ksn = nks - KPS * lc < KPS ? nks - KPS * lc : KPS;
            const uint32_t a_lo = desc_lo(a_base + s * STAGE_BYTES, 16);
            for (int ks = 0; ks < ksn; ++ks, ++kk) {
              const uint64_t ad = desc_join(a_lo + a_kinc * (uint32_t)ks, hi_a);
              if constexpr (X2) {
                const uint64_t ad_lo = desc_join(a_lo + a_kinc * (uint32_t)ks + 2u, hi_a);
#pragma unroll
                for (int term = 0; term < 3; ++term) {
#pragma unroll
                  for (int j = 0; j < TAPS; ++j) {
                    if (!valid[j]) continue;
                    umma_bf16_elect(dcol[j] + cvoff_d, term == 2 ? ad_lo : ad,
                                    desc_join(bdesc0[j] + cvoff_b + kk * wstep + (term == 1 ? (wtile >> 4) : 0u), hi_b), idesc,
                                    started[j] | ((kk > 0 || term > 0) ? 1u : 0u));
                  }
                }
                continue;
              }
#pragma unroll
              for (int j = 0; j < TAPS; ++j) {
                if (!valid[j]) continue;
                umma_bf16_elect(dcol[j] + cvoff_d, ad, desc_join(bdesc0[j] + cvoff_b + kk * wstep, hi_b), idesc,
                                started[j] | (kk > 0 ? 1u : 0u));
              }
            }
            umma_commit_elect(empty_bar(s));
            if (++s == NST) { s = 0; ph ^= 1u; }
          }
          // outputs completed by this frame
          if (TAPS == 3) {
            if (fi >= 1) umma_commit_elect(tfull_bar((int)((uint32_t)(gbase + fi - 1) & amask)));
            if (fi == T - 1) umma_commit_elect(tfull_bar((int)((uint32_t)(gbase + fi) & amask)));
          } else {
            umma_commit_elect(tfull_bar((int)((uint32_t)(gbase + fi) & amask)));
          }
        }
      }
      if (p.dbg && blockIdx.x == 0 && lane == 0) { p.dbg[2] = w_full; p.dbg[3] = w_tempty; p.dbg[4] = clock64() - t_start; p.dbg[5] = t_w; }
    }
  } else {
    // ===================== epilogue warps 2..5 =====================
    const int q = warp & 3;
    const int row = q * 32 + lane;
    pdl_wait();
    const size_t Mtot = (size_t)p.B * T * p.hw;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16);
    long long w_tfull = 0, w_cp = 0;
    const long long t_start_e = clock64();
    uint32_t gframe = 0;     // frames processed so far by this CTA: staging buffer = gframe & 1
    if (p.epi_quads) {
      stage_epilogue_operands(p, (int)blockIdx.x, 0, 0, row, Mtot, epi_base + (uint32_t)row * 16u);
      stage_epilogue_operands(p, (int)blockIdx.x, 0, 1, row, Mtot, epi_base + (uint32_t)(p.epi_quads * MT + row) * 16u);
    }
    int it = 0;
    for (int tile_i = blockIdx.x; tile_i < p.ntiles; tile_i += gridDim.x, ++it) {
      const int tile = p.zrev ? p.ntiles - 1 - tile_i : tile_i;
      const int b = tile / p.tiles_p;
      const int pix = (tile - b * p.tiles_p) * MT + row;
      const bool valid = pix < p.hw;
      for (int t = 0; t < T; ++t, ++gframe) {
        const size_t m = ((size_t)b * T + t) * p.hw + pix;
        const int aslot = (int)(gframe & (uint32_t)(NACC - 1));     // rolling accumulator of this output frame (NACC = 2 or 4)
        const uint32_t acol = lane_addr + (uint32_t)(aslot * npad);
        timed_wait(tfull_bar(aslot), (gframe / (uint32_t)NACC) & 1u, p.err, 15, w_tfull);
        tc_fence_after();
        const int ncols = p.epi == EPI_COUPLE_Y1 ? 16 : (p.epi == EPI_STORE ? ((p.cout + 15) & ~15) : kHF);
        const bool live = valid && (long long)m < p.m_limit;
        if (GMM) {                                    // its own instantiation: 48 softmax weights live in registers
          // ---- fused GMM head + sampler (SelfC_GMM_arch_inv.py:383-394): columns [0,48) logits, [48,96) log-scales,
          // [96,144) means of component k; softmax over the 48 hf channels
          float pi[kHF];
          float mx = -INFINITY;
#pragma unroll
          for (int c3 = 0; c3 < 3; ++c3) {
            uint32_t r[16];
            tmem_ld16(acol + (uint32_t)(16 * c3), r);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              pi[16 * c3 + j] = __uint_as_float(r[j]) + sbias[16 * c3 + j];
              mx = fmaxf(mx, pi[16 * c3 + j]);
            }
          }
          float sum = 0.f;
#pragma unroll
          for (int j = 0; j < kHF; ++j) {
            pi[j] = __expf(pi[j] - mx);
            sum += pi[j];
          }
          const float inv = 1.0f / sum;
          const size_t Mreal = (size_t)p.m_limit;
          const long long n = live ? (long long)m / p.gmm_hw : 0;
          const long long gpix = (long long)m - n * p.gmm_hw;
          const int gt = (int)(n % p.gmm_T);
          const long long gb = n / p.gmm_T;
#pragma unroll
          for (int c3 = 0; c3 < 3; ++c3) {
            uint32_t rl[16], rm[16];
            tmem_ld16(acol + (uint32_t)(kHF + 16 * c3), rl);
            tmem_ld16(acol + (uint32_t)(2 * kHF + 16 * c3), rm);
            tmem_ld_wait();
            if (c3 == 2) {                       // the whole accumulator is in registers: hand it back to the MMA warp
              tc_fence_before();
              __syncwarp();
              if (lane == 0) mbar_arrive(tempty_bar(aslot));
            }
            if (!live) continue;
#pragma unroll
            for (int q4 = 0; q4 < 4; ++q4) {
              const int hfq = 4 * c3 + q4;
              float ep4[4];
              if (p.eps) {
#pragma unroll
                for (int e = 0; e < 4; ++e)
                  ep4[e] = __ldg(p.eps + (uint64_t)(((((gb * kHF + 4 * hfq + e) * kGmmK + p.gmm_k) * p.gmm_T + gt) * p.gmm_hw) + gpix));
              } else {
                philox_normal4(eps_group(gb, hfq, p.gmm_k, gt, gpix, p.gmm_T, p.gmm_hw), p.seed, p.offset, ep4);
              }
              float o[4];
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const int hf = 4 * hfq + e, cc = 4 * q4 + e;
                const float ls = fminf(fmaxf(__uint_as_float(rl[cc]) + sbias[kHF + hf], -7.f), 7.f);
                const float mu = __uint_as_float(rm[cc]) + sbias[2 * kHF + hf];
                o[e] = pi[hf] * inv * fmaf(ep4[e], __expf(ls), mu);
              }
              float* zq = p.z + quad_off(Mreal, 1 + hfq, m);
              if (p.gmm_k > 0) {
                const float4 old = ld_nc4(zq);
                o[0] += old.x; o[1] += old.y; o[2] += old.z; o[3] += old.w;
              }
              store4(zq, make_float4(o[0], o[1], o[2], o[3]));
            }
          }
          continue;
        }
        const uint32_t ebuf = epi_base + (uint32_t)((gframe & 1) * p.epi_quads * MT + row) * 16u;
        if (p.epi_quads) {
          const long long t0 = clock64();
          asm volatile("cp.async.wait_group 1;" ::: "memory");   // this frame's operands have landed
          w_cp += clock64() - t0;
        }
        for (int n0 = 0; n0 < ncols; n0 += 16) {
          float4 x2q[4], sq[4];
          if (TAPS == 3 && live && p.epi == EPI_COUPLE_HG) {
#pragma unroll
            for (int j = 0; j < 4; ++j) x2q[j] = lds4(ebuf + (uint32_t)((n0 / 4 + j) * MT) * 16u);
          } else if (TAPS == 3 && live && p.epi == EPI_COUPLE_Y2) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              x2q[j] = lds4(ebuf + (uint32_t)((n0 / 4 + j) * MT) * 16u);
              sq[j] = lds4(ebuf + (uint32_t)((kSQuads + n0 / 4 + j) * MT) * 16u);
            }
          } else if (TAPS == 3 && live && p.epi == EPI_COUPLE_Y1) {
            x2q[0] = lds4(ebuf);
          }
          uint32_t r[16], rg[16];
          tmem_ld16(acol + (uint32_t)n0, r);
          if (TAPS == 3 && p.epi == EPI_COUPLE_HG) tmem_ld16(acol + (uint32_t)(kHF + n0), rg);     // warp-uniform: G's columns of the accumulator
          tmem_ld_wait();
          if (!live) continue;
          float v[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(r[j]) + sbias[n0 + j];

          switch (TAPS == 1 ? EPI_STORE : p.epi) {     // the pointwise instantiation only stores (the coupling epilogues belong to conv5)
            case EPI_STORE: {
              if (p.act) {
#pragma unroll
                for (int j = 0; j < 16; ++j) v[j] = lrelu02(v[j]);
              }
              if (X2 && p.outT) {       // cout % 16 == 0 (checked by the launcher)
                __nv_bfloat16* o = p.outT + x2_hi_index(dense_off((long long)m, p.outT_off + n0, p.outT_pitch, p.outT_slabM));
                x2_store8(o, v);
                x2_store8(o + 8, v + 8);
              } else if (p.outT) {
                __nv_bfloat16* o = p.outT + dense_off((long long)m, p.outT_off + n0, p.outT_pitch, p.outT_slabM);
                if (n0 + 16 <= p.cout) {
                  store_bf16x8(o, v);
                  store_bf16x8(o + 8, v + 8);
                } else {
                  for (int j = 0; j < 16; ++j)
                    if (n0 + j < p.cout) o[j] = __float2bfloat16_rn(v[j]);
                }
              }
              const int ncol = p.outF_blk ? p.outF_off + (n0 / p.outF_blk) * p.outF_blk_stride + n0 % p.outF_blk : p.outF_off + n0;
              if (p.outF && p.outF_planar == 2) {
                __half* oh = reinterpret_cast<__half*>(p.outF);
#pragma unroll
                for (int j = 0; j < 16; j += 4) {
                  const __half2 a = __floats2half2_rn(v[j], v[j + 1]), b = __floats2half2_rn(v[j + 2], v[j + 3]);
                  uint2 pk;
                  pk.x = *reinterpret_cast<const uint32_t*>(&a);
                  pk.y = *reinterpret_cast<const uint32_t*>(&b);
                  *reinterpret_cast<uint2*>(oh + quad_off((size_t)p.m_limit, (ncol + j) / 4, m)) = pk;
                }
              } else if (p.outF && p.outF_planar) {
#pragma unroll
                for (int j = 0; j < 16; j += 4)
                  store4(p.outF + quad_off((size_t)p.m_limit, (ncol + j) / 4, m), make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]));
              } else if (p.outF) {
                float* o = p.outF + dense_off((long long)m, p.outF_off + n0, p.outF_pitch, p.outF_slabM);
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
              float* zp = p.z + quad_off(Mtot, 0, m);
              const float4 x1 = x2q[0];
              float y[16];
#pragma unroll
              for (int j = 0; j < 16; ++j) y[j] = 0.f;
              y[0] = p.rev ? x1.x - v[0] : x1.x + v[0];
              y[1] = p.rev ? x1.y - v[1] : x1.y + v[1];
              y[2] = p.rev ? x1.z - v[2] : x1.z + v[2];
              store4(zp, make_float4(y[0], y[1], y[2], 0.f));
              if (X2) {
                if (p.copyA) {
                  __nv_bfloat16* o = p.copyA + x2_hi_index(dense_off((long long)m, 0, p.copyA_pitch, p.copy_slabM));
                  x2_store8(o, y);
                  x2_store8(o + 8, y + 8);
                }
                if (p.copyB) {
                  __nv_bfloat16* o = p.copyB + x2_hi_index(dense_off((long long)m, 0, p.copyB_pitch, p.copy_slabM));
                  x2_store8(o, y);
                  x2_store8(o + 8, y + 8);
                }
                break;
              }
              if (p.copyA) {
                __nv_bfloat16* o = p.copyA + dense_off((long long)m, 0, p.copyA_pitch, p.copy_slabM);
                store_bf16x8(o, y);
                if (p.copy_pad > 8) store_bf16x8(o + 8, y + 8);
              }
              if (p.copyB) {
                __nv_bfloat16* o = p.copyB + dense_off((long long)m, 0, p.copyB_pitch, p.copy_slabM);
                store_bf16x8(o, y);
                if (p.copy_pad > 8) store_bf16x8(o + 8, y + 8);
              }
            } break;
            case EPI_COUPLE_HG: {
              // columns [n0, n0+16) hold H (v = h + bias), columns [48+n0, ...) hold G: s = 2*sigmoid(h)-1 stays in registers
              float y[16];
#pragma unroll
              for (int j = 0; j < 16; j += 4) {
                float* zp = p.z + quad_off(Mtot, 1 + (n0 + j) / 4, m) - j;
                const float4 x2 = x2q[j / 4];
                const float xr[4] = {x2.x, x2.y, x2.z, x2.w};
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  const float sv = __fdividef(2.0f, 1.0f + __expf(-v[j + e])) - 1.0f;
                  const float g = __uint_as_float(rg[j + e]) + sbias[kHF + n0 + j + e];
                  y[j + e] = p.rev ? (xr[e] - g) * __expf(-sv) : fmaf(xr[e], __expf(sv), g);
                }
                store4(zp + j, make_float4(y[j], y[j + 1], y[j + 2], y[j + 3]));
              }
              if (p.copyA) {
                __nv_bfloat16* o = p.copyA + dense_off((long long)m, n0, p.copyA_pitch, p.copy_slabM);
                store_bf16x8(o, y);
                store_bf16x8(o + 8, y + 8);
              }
            } break;
            case EPI_COUPLE_S: {
#pragma unroll
              for (int j = 0; j < 16; j += 4) {
                float s[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) s[e] = __fdividef(2.0f, 1.0f + __expf(-v[j + e])) - 1.0f;   // 2*sigmoid(h)-1, ~1e-6 rel
                store4(p.sbuf + quad_off(Mtot, (n0 + j) / 4, m), make_float4(s[0], s[1], s[2], s[3]));
              }
            } break;
            case EPI_COUPLE_Y2: {
              float y[16];
#pragma unroll
              for (int j = 0; j < 16; j += 4) {
                float* zp = p.z + quad_off(Mtot, 1 + (n0 + j) / 4, m) - j;
                const float4 x2 = x2q[j / 4];
                const float4 s4 = sq[j / 4];
                const float xr[4] = {x2.x, x2.y, x2.z, x2.w};
                const float sr[4] = {s4.x, s4.y, s4.z, s4.w};
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  // s in (-1,1): exp via ex2.approx (rel 2^-21); the reverse divides by multiplying with exp(-s)
                  y[j + e] = p.rev ? (xr[e] - v[j + e]) * __expf(-sr[e]) : fmaf(xr[e], __expf(sr[e]), v[j + e]);
                }
                store4(zp + j, make_float4(y[j], y[j + 1], y[j + 2], y[j + 3]));
              }
              if (X2 && p.copyA) {
                __nv_bfloat16* o = p.copyA + x2_hi_index(dense_off((long long)m, n0, p.copyA_pitch, p.copy_slabM));
                x2_store8(o, y);
                x2_store8(o + 8, y + 8);
              } else if (p.copyA) {
                __nv_bfloat16* o = p.copyA + dense_off((long long)m, n0, p.copyA_pitch, p.copy_slabM);
                store_bf16x8(o, y);
                store_bf16x8(o + 8, y + 8);
              }
            } break;
            default: break;
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(tempty_bar(aslot));
        if (p.epi_quads) stage_epilogue_operands(p, tile_i, t, 2, row, Mtot, ebuf);
      }
    }
    if (p.epi_quads) asm volatile("cp.async.wait_group 0;" ::: "memory");
    if (p.dbg && blockIdx.x == 0 && threadIdx.x == 64) { p.dbg[6] = w_tfull; p.dbg[7] = w_cp; p.dbg[8] = clock64() - t_start_e; p.dbg[9] = it; }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
  }
}

// wref [cout][cin_ref][taps] fp32 -> bf16 image [tap][kstep][kcore(2)][ngroup(npad/8)][n%8][k%8]
// x2 (BF16X3 mode): [tap][kstep][hi|lo][kcore(2)][ngroup(npad/8)][n%8][k%8], hi = bf16(w), lo = bf16(w - hi)
struct PackTemporalJob {
  const float* wref;
  const float* bref;
  __nv_bfloat16* img;
  float* bias;
  int cout, cin_ref, taps, cin_buf, xreal, xpad, npad, x2;
};

__device__ __forceinline__ void pack_temporal_body(const PackTemporalJob& j, int idx) {
  const float* __restrict__ wref = j.wref;
  const float* __restrict__ bref = j.bref;
  __nv_bfloat16* __restrict__ img = j.img;
  float* __restrict__ bias = j.bias;
  const int cout = j.cout, cin_ref = j.cin_ref, taps = j.taps, cin_buf = j.cin_buf, xreal = j.xreal, xpad = j.xpad, npad = j.npad, x2 = j.x2;
  const int nchunk = cin_buf / 16;
  const int total = taps * cin_buf * npad;
  if (idx < npad) bias[idx] = idx < cout ? bref[idx] : 0.f;
  if (idx >= total) return;
  const int n = idx % npad;
  const int c = (idx / npad) % cin_buf;
  const int tap = idx / (npad * cin_buf);
  const int cref = c < xreal ? c : (c < xpad ? -1 : c - xpad + xreal);
  float v = 0.f;
  if (n < cout && cref >= 0 && cref < cin_ref) v = wref[((size_t)n * cin_ref + cref) * taps + tap];
  const int ks = c / 16, kk = c % 16;
  const size_t inner = (size_t)((kk / 8) * (npad / 8) + n / 8) * 64 + (n % 8) * 8 + (kk % 8);
  if (x2) {
    const size_t off = (size_t)(tap * nchunk + ks) * 2 * (npad * 16) + inner;
    __nv_bfloat16 hi, lo;
    x2_split(v, hi, lo);
    img[off] = hi;
    img[off + (size_t)npad * 16] = lo;
    return;
  }
  const size_t off = (size_t)(tap * nchunk + ks) * (npad * 16) + inner;
  img[off] = __float2bfloat16_rn(v);
}

__global__ void pack_temporal_kernel(const PackTemporalJob j) { pack_temporal_body(j, blockIdx.x * blockDim.x + threadIdx.x); }

// every recorded job in one launch (pack_batch.h)
__global__ void pack_temporal_multi_kernel(const PackTemporalJob* __restrict__ jobs, const int* __restrict__ first, int njobs) {
  const int ji = pack_find_job(first, njobs, blockIdx.x);
  const PackTemporalJob j = jobs[ji];
  pack_temporal_body(j, (blockIdx.x - __ldg(first + ji)) * blockDim.x + threadIdx.x);
}

}  // namespace tc5

int pack_temporal_weights(TcTempW& w, const float* wref, const float* bref, int cout, int cin_ref, int taps, int cin_buf, int xreal,
                          int xpad, cudaStream_t st, bool x2) {
  SELFC_CHECK_ARG(cin_buf % 16 == 0 && cout >= 1 && cout <= 256, "temporal_tc: cin %d / cout %d unsupported", cin_buf, cout);
  const int npad = (cout + 15) & ~15;
  const size_t bytes = (size_t)taps * (cin_buf / 16) * npad * 32 * (x2 ? 2 : 1);
  if (w.img == nullptr || w.img_bytes != bytes) {
    free_temporal_weights(w);
    SELFC_CUDA(cudaMalloc(&w.img, bytes));
    SELFC_CUDA(cudaMalloc(&w.bias, 256 * sizeof(float)));
    w.img_bytes = bytes;
  }
  w.cin_buf = cin_buf; w.npad = npad; w.taps = taps; w.cout = cout; w.x2 = x2;
  const int total = taps * cin_buf * npad;
  const tc5::PackTemporalJob j{wref, bref, reinterpret_cast<__nv_bfloat16*>(w.img), w.bias, cout, cin_ref, taps, cin_buf, xreal, xpad, npad, x2 ? 1 : 0};
  if (PackBatch* pb = pack_batch_current()) {
    pb->temporal[pb->point].add(j, (int)cdiv(total, 256));
    return 0;
  }
  tc5::pack_temporal_kernel<<<cdiv(total, 256), 256, 0, st>>>(j);
  SELFC_LAUNCH_CHECK("pack_temporal_kernel");
  return 0;
}

int flush_pack_temporal(JobTable& t, cudaStream_t st) {
  if (t.njobs() <= 0) return 0;
  const void* jobs = nullptr;
  const int* first = nullptr;
  SELFC_CUDA(t.sync(st, &jobs, &first));
  tc5::pack_temporal_multi_kernel<<<t.first.back(), 256, 0, st>>>(static_cast<const tc5::PackTemporalJob*>(jobs), first, t.njobs());
  SELFC_LAUNCH_CHECK("pack_temporal_multi_kernel");
  return 0;
}

void free_temporal_weights(TcTempW& w) {
  if (w.img) cudaFree(w.img);
  if (w.bias) cudaFree(w.bias);
  w.img = nullptr;
  w.bias = nullptr;
  w.img_bytes = 0;
}

bool temporal_tc_supported(const TcTempW& w, int T) { return w.img != nullptr && T >= 1 && T <= 32; }

int launch_temporal_tc(const TcTempW& w, const TcTempArgs& a, cudaStream_t st, const TcTempW* w2) {
  SELFC_CHECK_ARG(temporal_tc_supported(w, a.T), "temporal_tc: weights not packed or T=%d outside [1,32]", a.T);
  const bool hg = a.epi == EPI_COUPLE_HG;
  SELFC_CHECK_ARG(!hg || (w2 != nullptr && w2->img != nullptr && w2->cin_buf == w.cin_buf && w2->npad == w.npad && w.npad == kHF &&
                          w.taps == 3 && a.in2 != nullptr && a.in_slabM != 0 && aligned16(a.in2)),
                  "temporal_tc: the fused H+G epilogue needs two 48-output temporal convs over slab-planar buffers of the same shape");
  const bool x2 = w.x2;
  SELFC_CHECK_ARG(!x2 || (!hg && a.epi != EPI_GMM && a.epi != EPI_COUPLE_HG && ((uintptr_t)a.in & 63) == 0 && a.in_pitch % 16 == 0 &&
                          (a.outT == nullptr || (w.cout % 16 == 0 && a.outT_off % 16 == 0 && a.outT_pitch % 16 == 0))),
                  "temporal_tc: the (hi, lo) form has no fused H+G / GMM epilogue and needs 64-byte aligned rows");
  const int npad_acc = hg ? 2 * w.npad : w.npad;       // accumulator columns per output frame
  SELFC_CHECK_ARG(a.in_pitch % 8 == 0 && aligned16(a.in), "temporal_tc: input pitch/alignment");
  SELFC_CHECK_ARG(a.epi != EPI_GA, "temporal_tc: the GlobalAgg mix is a separate kernel (stp.cu: ga_mix)");
  tc::EncodeTiledFn encode = tc::get_encode_fn();
  if (!encode) {
    set_error("cuTensorMapEncodeTiled is not available from this driver");
    return SELFC_E_CUDA;
  }
  const int BT = a.B * a.T;
  const int nks = w.cin_buf / 16;
  // K steps per ring stage: slab input -> the split of nks with the least padding (a stage is kps sub-tiles of 4 KB);
  // pixel-major input -> one SWIZZLE_128B row of 64 channels per pixel
  int kps = x2 ? 2 : 4;
  if (a.in_slabM) {
    int best_waste = 1 << 30;
    for (int k = 4; k >= 2; --k) {
      const int waste = cdiv(nks, k) * k - nks;
      if (waste < best_waste) { best_waste = waste; kps = k; }
    }
    if (nks < kps) kps = nks;
    if (x2) {
      // (hi, lo) stages are twice as large and the weights twice as many: keep at least three ring slots
      const int epi_q = a.epi == EPI_COUPLE_Y2 ? 2 * kSQuads : (a.epi == EPI_COUPLE_Y1 ? 1 : 0);
      const int fixed0 = tc5::BAR_BYTES + tc5::BIAS_BYTES + (int)w.img_bytes + 2 * epi_q * tc5::MT * 16 + 1024;
      while (kps > 1 && (227 * 1024 - fixed0) / (kps * tc5::MT * 64) < 3) --kps;
    }
  }
  if (hg && nks % kps != 0) return SELFC_E_UNSUPPORTED;      // the second conv's chunks must start on a stage boundary
  const int nchunk1 = cdiv(nks, kps);
  const int nchunk = hg ? 2 * nchunk1 : nchunk1;
  CUtensorMap tmap, tmap2;
  CUresult r;
  if (a.in_slabM) {
    // slab-planar dense buffer [cin/16][M][16]: box = 128 pixels of one frame x kps slabs, each slab's rows one 4 KB run
    SELFC_CHECK_ARG(a.in_slabM == (long long)BT * a.hw, "temporal_tc: slab stride %lld != B*T*hw", a.in_slabM);
    const cuuint64_t rowb = x2 ? 64 : 32;      // bytes per (pixel, slab) row; x2: [16 x hi | 16 x lo]
    const cuuint64_t gdim[4] = {rowb / 2, (cuuint64_t)a.hw, (cuuint64_t)BT, (cuuint64_t)nks};
    const cuuint64_t gstr[3] = {rowb, (cuuint64_t)a.hw * rowb, (cuuint64_t)a.in_slabM * rowb};
    const cuuint32_t box[4] = {(cuuint32_t)(rowb / 2), (cuuint32_t)tc5::MT, 1, (cuuint32_t)kps};
    const cuuint32_t estr[4] = {1, 1, 1, 1};
    const CUtensorMapSwizzle swz = x2 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B;
    r = encode(&tmap, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<__nv_bfloat16*>(a.in), gdim, gstr, box, estr,
               CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r == CUDA_SUCCESS)
      r = encode(&tmap2, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<__nv_bfloat16*>(hg ? a.in2 : a.in), gdim, gstr, box, estr,
                 CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                 CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  } else {
    // x2: a pixel's row is in_pitch (hi, lo) pairs = 2 * in_pitch bf16; a stage still takes 64 bf16 = 2 K-steps of [16 hi | 16 lo]
    const int epp = x2 ? 2 * a.in_pitch : a.in_pitch;
    const cuuint64_t gdim[3] = {(cuuint64_t)epp, (cuuint64_t)a.hw, (cuuint64_t)BT};
    const cuuint64_t gstr[2] = {(cuuint64_t)epp * 2, (cuuint64_t)a.hw * epp * 2};
    const cuuint32_t box[3] = {64, (cuuint32_t)tc5::MT, 1};
    const cuuint32_t estr[3] = {1, 1, 1};
    r = encode(&tmap, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<__nv_bfloat16*>(a.in), gdim, gstr, box, estr,
               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    tmap2 = tmap;
  }
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled (temporal) failed with CUresult %d", (int)r);
    return SELFC_E_CUDA;
  }
  tc5::Params p;
  memset(&p, 0, sizeof(p));
  p.wimg = w.img; p.bias = w.bias;
  p.wimg2 = hg ? w2->img : w.img; p.bias2 = hg ? w2->bias : w.bias;
  p.T = a.T; p.B = a.B; p.hw = a.hw; p.nks = nks; p.npad = npad_acc; p.taps = w.taps; p.cout = w.cout;
  p.nhalf = w.npad;
  p.kps = kps; p.nchunk = nchunk; p.nc_split = nchunk1;
  p.tiles_p = cdiv(a.hw, tc5::MT);
  p.ntiles = p.tiles_p * a.B;
  // taps == 3: the outputs fi-1, fi, fi+1 are being accumulated while the epilogue drains a fourth one
  const int nacc = 512 / npad_acc >= 4 ? 4 : (512 / npad_acc >= 2 ? 2 : 0);     // power of two: slot = index & (nacc-1)
  SELFC_CHECK_ARG(w.taps == 3 || w.taps == 1, "temporal_tc: %d taps", w.taps);
  if (nacc < (w.taps == 3 ? 4 : 2)) {
    set_error("temporal_tc: N=%d leaves no room for %d rolling accumulators", npad_acc, w.taps == 3 ? 4 : 2);
    return SELFC_E_UNSUPPORTED;
  }
  p.nacc = nacc;
  int cols = nacc * npad_acc, pw = 32;
  while (pw < cols) pw <<= 1;
  p.tmem_cols = pw;
  p.epi = a.epi; p.rev = a.rev; p.act = a.act;
  p.in_slab = a.in_slabM ? 1 : 0;
  p.outT = a.outT; p.outT_pitch = a.outT_pitch; p.outT_off = a.outT_off;
  p.outT_slabM = a.outT_slabM; p.copy_slabM = a.copy_slabM;
  p.outF = a.outF; p.outF_pitch = a.outF_pitch; p.outF_off = a.outF_off; p.outF_planar = a.outF_planar;
  p.outF_blk = a.outF_blk; p.outF_blk_stride = a.outF_blk_stride; p.outF_slabM = a.outF_slabM;
  SELFC_CHECK_ARG(a.outF_slabM == 0 || (a.outF_off % 16 == 0 && a.outF_pitch % 16 == 0 && w.cout % 16 == 0), "temporal_tc: slab-planar fp32 output");
  p.m_limit = a.m_limit > 0 ? a.m_limit : (long long)BT * a.hw;
  p.z = a.z; p.sbuf = a.sbuf;
  p.copyA = a.copyA; p.copyA_pitch = a.copyA_pitch; p.copyB = a.copyB; p.copyB_pitch = a.copyB_pitch; p.copy_pad = a.copy_pad;
  p.wmat = a.wmat; p.wsum = a.wsum; p.resid = a.resid; p.resid_pitch = a.resid_pitch;
  p.outAct = a.outAct; p.outAct_pitch = a.outAct_pitch;
  p.gmm_k = a.gmm_k; p.gmm_T = a.gmm_T; p.gmm_hw = a.gmm_hw; p.eps = a.eps; p.seed = a.seed; p.offset = a.offset;
  SELFC_CHECK_ARG(a.epi != EPI_GMM || (w.npad == 3 * kHF && a.z != nullptr && a.gmm_hw > 0 && a.gmm_T >= 1 && a.gmm_k >= 0 && a.gmm_k < kGmmK),
                  "temporal_tc: GMM epilogue needs N = 144, the latent state and the clip shape");
  p.err = tc::err_flag_for_device();
  p.dbg = reinterpret_cast<long long*>(a.dbg);
  if (!p.dbg && tc::debug_slots()) {          // SELFC_TC_DBG=1: every launch records CTA 0's barrier-wait cycles
    long long* slot = tc::debug_next_slot(a.epi * 1000000 + w.npad * 1000 + w.cin_buf / 16);
    p.dbg = slot;
  }
  if (p.ntiles == 0) return 0;
  p.zrev = tc::next_direction();
  p.epi_quads = a.epi == EPI_COUPLE_Y2 ? 2 * kSQuads : (a.epi == EPI_COUPLE_HG ? kSQuads : (a.epi == EPI_COUPLE_Y1 ? 1 : 0));
  const int fixed = tc5::BAR_BYTES + tc5::BIAS_BYTES + (int)w.img_bytes * (hg ? 2 : 1) + 2 * p.epi_quads * tc5::MT * 16 + 1024;
  const int stage_bytes = kps * tc5::MT * (x2 ? 64 : 32);
  int nst = (227 * 1024 - fixed) / stage_bytes;
  if (nst > tc5::NST_MAX) nst = tc5::NST_MAX;
  if (nst < 2) {
    set_error("temporal_tc: no room for two %d-byte stages next to %zu bytes of weights", stage_bytes, w.img_bytes);
    return SELFC_E_UNSUPPORTED;
  }
  p.nst = nst;
  const int smem = nst * stage_bytes + fixed;
  static bool smem_set[64] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev >= 0 && dev < 64 && !smem_set[dev]) {
    SELFC_CUDA(cudaFuncSetAttribute(tc5::temporal_tc_kernel<3, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    SELFC_CUDA(cudaFuncSetAttribute(tc5::temporal_tc_kernel<1, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    SELFC_CUDA(cudaFuncSetAttribute(tc5::temporal_tc_kernel<1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    SELFC_CUDA(cudaFuncSetAttribute(tc5::temporal_tc_kernel<3, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    SELFC_CUDA(cudaFuncSetAttribute(tc5::temporal_tc_kernel<1, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    smem_set[dev] = true;
  }
  const int nsm = tc::num_sms();
  const int grid = p.ntiles < nsm ? p.ntiles : nsm;
  SELFC_CHECK_ARG(a.epi != EPI_GMM || w.taps == 1, "temporal_tc: the GMM epilogue belongs to a pointwise conv");
  if (x2 && w.taps == 3) SELFC_CUDA(tc::launch_pdl(tc5::temporal_tc_kernel<3, false, true>, grid, tc5::THREADS, smem, st, tmap, tmap2, p));
  else if (x2) SELFC_CUDA(tc::launch_pdl(tc5::temporal_tc_kernel<1, false, true>, grid, tc5::THREADS, smem, st, tmap, tmap2, p));
  else if (w.taps == 3) SELFC_CUDA(tc::launch_pdl(tc5::temporal_tc_kernel<3, false>, grid, tc5::THREADS, smem, st, tmap, tmap2, p));
  else if (a.epi == EPI_GMM) SELFC_CUDA(tc::launch_pdl(tc5::temporal_tc_kernel<1, true>, grid, tc5::THREADS, smem, st, tmap, tmap2, p));
  else SELFC_CUDA(tc::launch_pdl(tc5::temporal_tc_kernel<1, false>, grid, tc5::THREADS, smem, st, tmap, tmap2, p));
  SELFC_LAUNCH_CHECK("temporal_tc_kernel");
  return 0;
}

}  // namespace selfc
