// (1,3,3) dense-block convolution conv1..4 of D2DTInput (Subnet_constructor.py:102-105,126-129) as a tcgen05 / TMEM /
// TMA implicit GEMM over the slab-planar dense buffer (common.cuh), BF16 mode.
//
// Formulation: PIXELS are the M dimension (one TMEM lane = one pixel, so the epilogue is thread-per-pixel), the three
// kx taps are STACKED IN N:
//     D[p][kx*32 + n] += sum_c X[p + ky*32][c] * W[ky,kx][n][c]        M = 128 halo positions, N = 96, K = 16
// for ky = 0..2, where p runs over the flattened (rows x 32) halo tile; the ky shift is a start-address offset of 32
// rows in the A descriptor, the kx shift is applied when the accumulator is read back:
//     out[r][x][n] = sum_kx D[(r, x + kx)][kx*32 + n]                     -> two warp shuffles (lane = x)
// So one activation tile load feeds all nine taps, the tile is read from shared memory once per ky (3x, not 9x), and an
// MMA moves 4 KB (A) + 3 KB (B) of shared memory for 48 cycles of math (the N = 32 formulation moved 5 KB for 16).
//
// Data movement: a pipeline stage is `kps` K-steps; ONE 5-D TMA box brings the (8+2) x 32 halo of an 8 x 30 output tile, every
// tile row of a K-step being one contiguous 1 KB run of the slab (out-of-image rows / columns / slabs are zero-filled by the
// TMA = the conv padding).  DRAM traffic is algorithmic: the conv reads exactly the cin/16 slabs it consumes.  Two forms of the
// box (template P2): {16 ch, 32 x, 10 y} rows of one position (SWIZZLE_32B), or -- default for even widths -- rows of two
// adjacent positions {32, 16 pairs, 10 y} (SWIZZLE_64B), which halves the number of rows the TMA unit has to move.
//
// Per CTA (persistent, 192 threads): warp 0 TMA producer, warp 1 MMA issuer (whole warp, elected lane), warps 2..5
// epilogue: tcgen05.ld -> shuffle-add the kx partials -> + bias -> LeakyReLU(0.2) -> bf16 -> 32-byte stores into the output
// slabs.  Accumulators are double-buffered across tiles (2 tiles x 2 accumulators x 128 columns), the whole conv's weights stay
// resident in shared memory.  By default CTAs run as pairs (template PAIR, cta_group::2): one M = 256 MMA covers the tiles of
// both CTAs and each CTA keeps only half of the weight rows.  What each form buys and why: profiles/r2_conv3x3_waits.md.
#include <cuda.h>
#include <stdlib.h>
#include <string.h>

#include "common.cuh"
#include "conv_tc.h"
#include "pack_batch.h"
#include "tc_ptx.cuh"

namespace selfc {
namespace tc3 {

using namespace tc;

constexpr int WT = 32;                        // tile pitch (30 valid columns + 2 halo)
constexpr int VALID_W = WT - 2;
constexpr int ROWS = 8;                       // output rows per tile = 2 M-blocks of 4 rows x 32 positions
constexpr int HT = ROWS + 2;
constexpr int HALO_POS = HT * WT;             // 320 halo positions = rows of the activation tile
constexpr int MBLK = 2;
constexpr int NOUT = 32;
constexpr int NB = 3 * NOUT;                  // MMA N: kx-major, 96
constexpr int SUB_BYTES = HALO_POS * 32;      // one K-step sub-tile: 320 rows x 16 bf16, SWIZZLE_32B
constexpr int WTILE_BYTES = NB * 16 * 2;      // one (ky, K-step) B tile: 96 x 16 bf16 = 3 KB
constexpr int ACC_STRIDE = 128;               // TMEM columns reserved per accumulator (96 used)
constexpr int TMEM_COLS = 512;                // 2 tiles x 2 M-blocks x 128
constexpr int MAX_CIN = 160;
constexpr int THREADS = 192;
constexpr int NSTAGE_MAX = 8;
constexpr int BAR_BYTES = 256;

struct Params {
  // up to two independent problems of identical shape in one launch (G and H of an InvBlockExp read the same y1 and
  // differ only in weights and buffer): CTA parity selects the problem, each gets half of the grid
  const void* wimg;
  const void* wimg2;
  const float* bias;
  const float* bias2;
  __nv_bfloat16* buf;
  __nv_bfloat16* buf2;
  int nprob;             // 1 or 2
  long long slabM;       // pixels per slab (= N*h*w)
  int out_slab, N, h, w, nks, kps;
  int tiles_x, tiles_y, ntiles;
  int nst;
  int rev;               // walk the tiles last-to-first (see tc::next_direction)
  // X2 input-gradient launches (training, BF16X3 mode): no bias / activation, the 32 results are ADDED to channels
  // [acc_off, acc_off + acc_n) of an fp32 pixel-major buffer instead of being stored as (hi, lo) pairs
  // Such a launch produces acc_n channels in groups of 32: `ngroups` weight images resident, every tile visited once per group.
  float* accF;
  int acc_pitch, acc_off, acc_n, ngroups;
  long long acc_slabM;   // layout of accF: 0 pixel-major, else 16-channel fp32 slabs of acc_slabM pixels
  int* err;
  long long* dbg;        // SELFC_TC_DBG=1 (+ -DSELFC_TC_TIMING): CTA 0's barrier-wait cycles
};

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

__device__ __forceinline__ void tma_load_5d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2, int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}

__device__ __forceinline__ uint32_t pack_bf2(float a, float b) {
  __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&v);
}

// Tile order: band-major -- (tile row, frame, tile column) from slowest to fastest -- so that a launch sweeps the image top to
// bottom across all frames, the order the (3,1,1) conv5 kernel also walks the pixels in; p.rev walks it bottom to top.
// Consecutive launches of a dense block alternate direction (tc::next_direction): a launch starts on the rows the previous
// one touched last, which are still in L2 (the dense buffer of a GOP is 2-4x the L2, so a same-direction sweep always misses).
__device__ __forceinline__ void decode_tile(const Params& p, int tile, int& tx, int& ty, int& n) {
  if (tile >= p.ntiles) { tx = 0; ty = 0; n = p.N; return; }
  const int t = p.rev ? p.ntiles - 1 - tile : tile;
  tx = t % p.tiles_x;
  n = (t / p.tiles_x) % p.N;
  ty = t / (p.tiles_x * p.N);
}

// PAIR: the kernel runs as CTA pairs (cluster of 2, cta_group::2).  Each CTA of a pair works on its own tile with its own
// activation pipeline, accumulators and epilogue; the pair's leader issues ONE M = 256 MMA for both (128 halo positions from
// each CTA), and each CTA keeps only half of the weight rows (48 of the 96 (kx, n) rows), so an MMA reads 4 + 1.5 KB of each
// SM's shared memory instead of 4 + 3 KB -- the operand feed, not the tensor pipe, is what bounds the one-CTA form.
// P2: the activation tile is fetched as rows of TWO adjacent positions (64-byte rows, SWIZZLE_64B) instead of one: the TMA unit
// delivers about one box row per clock whatever its width (scripts/ubench/tma_rate.cu), so this halves the TMA time per K-step
// (320 -> 160 rows for six MMAs), which is what bounds the mainloop of the one-position form.  An M-row of the MMA is then a
// position PAIR g and the K-step of the first / second position of the pair is selected by a +0 / +32 byte start address inside
// the 64-byte row, exactly like a K advance inside a swizzle atom: accumulator i (0/1) row g <-> position 2g+i, one MMA covers all
// 8 tile rows (128 pairs), the ky tap is a 16-pair (1 KB = two atoms) start offset.  Pairs are (x odd, x+1) so that the halo
// origin 30*tx-1 is pair aligned: the tensor map is based one position BEFORE the buffer, row pitch W positions, W/2+1 pairs per
// row.  The two positions this makes readable that are not padding zeros -- x = -1 (previous row's last pixel) and x = W (next
// row's first) -- only ever feed the kx=0 term of x = 0 and the kx=2 term of x = W-1, which the epilogue drops.
// X2 (BF16X3 mode, common.cuh; PAIR only, one position per TMA row): activations and weights are (hi, lo) bf16 pairs.  A position's
// 16 channels are one 64-byte row [16 x hi | 16 x lo] (SWIZZLE_64B), i.e. two K = 16 operand rows 32 bytes apart, every (ky, K-step)
// has a hi and a lo weight tile, and a tap issues THREE MMAs into the same accumulator -- hi.hi, hi.lo, lo.hi -- so the product
// carries 16 mantissa bits per operand with fp32 accumulation; the epilogue splits its fp32 results the same way.
template <bool PAIR, bool P2, bool X2 = false>
__global__ void __launch_bounds__(THREADS, 1) conv3x3_tc3_kernel(const __grid_constant__ CUtensorMap tmap_a,
                                                                  const __grid_constant__ CUtensorMap tmap_b, const Params p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  static_assert(!X2 || (PAIR && !P2), "the (hi, lo) form exists for CTA pairs with one position per TMA row");
  constexpr int NBH = PAIR ? NB / 2 : NB;                 // weight rows held by this CTA
  constexpr int WT_BYTES = NBH * 16 * 2;                  // one (ky, K-step) B tile of this CTA
  constexpr int WSTEP = X2 ? 2 * WT_BYTES : WT_BYTES;     // X2: a K-step's hi tile followed by its lo tile
  constexpr int SUBB = X2 ? 2 * SUB_BYTES : SUB_BYTES;    // one K-step sub-tile of activations
  const uint32_t crank = PAIR ? cluster_ctarank() : 0u;   // 0 = leader
  const int unit = PAIR ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;          // scheduling unit: a CTA or a CTA pair
  const int nunits = PAIR ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  const int prob = p.nprob == 2 ? (unit & 1) : 0;
  const int rank = p.nprob == 2 ? (unit >> 1) : unit;                        // this unit walks steps rank, rank + nwalk, ...
  const int nwalk = p.nprob == 2 ? ((nunits + 1 - prob) >> 1) : nunits;
  const int nsteps = PAIR ? (p.ntiles + 1) >> 1 : p.ntiles;                  // a pair takes tiles 2*step, 2*step + 1
  const CUtensorMap* tmap = prob ? &tmap_b : &tmap_a;
  const void* wimg = prob ? p.wimg2 : p.wimg;
  const float* biasp = prob ? p.bias2 : p.bias;
  __nv_bfloat16* obuf = prob ? p.buf2 : p.buf;
  const int NST = p.nst;
  const int KPS = p.kps;
  const int STAGE = KPS * SUBB;
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gen_base = smem_raw + (base - smem_u32(smem_raw));
  // [activation stages][barriers][weights]
  const uint32_t a_base = base;
  const int bar_off = NST * STAGE;
  const uint32_t bar_base = base + bar_off;
  const uint32_t w_base = base + bar_off + BAR_BYTES;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (NSTAGE_MAX + s); };
  const uint32_t w_bar = bar_base + 8u * (2 * NSTAGE_MAX);
  auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * NSTAGE_MAX + 1 + a); };
  auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * NSTAGE_MAX + 3 + a); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * NSTAGE_MAX + 5);
  const uint32_t wpeer_bar = bar_base + 8u * (2 * NSTAGE_MAX + 6);           // leader: the peer's weights have landed
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(gen_base + bar_off + 8 * (2 * NSTAGE_MAX + 5));

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < NSTAGE_MAX; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    mbar_init(w_bar, 1);
    mbar_init(wpeer_bar, 1);
    for (int a = 0; a < 2; ++a) {
      mbar_init(tfull_bar(a), 1);
      mbar_init(tempty_bar(a), PAIR ? 8 : 4);       // the epilogue warps of both CTAs release the leader's accumulator stage
    }
    fence_barrier_init();
    asm volatile("prefetch.tensormap [%0];" ::"l"(tmap) : "memory");
  }
  if (warp == 1) {
    if (PAIR) tmem_alloc2(tmem_slot, (uint32_t)TMEM_COLS);
    else tmem_alloc(tmem_slot, (uint32_t)TMEM_COLS);
  }
  tc_fence_before();
  __syncthreads();
  if (PAIR) cluster_sync_all();      // the peer's barriers exist before anything is signalled across the pair
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  pdl_launch_dependents();

  const int nks = p.nks;
  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      const uint32_t wbytes = 3u * nks * WSTEP;
      mbar_expect_tx(w_bar, wbytes * (uint32_t)p.ngroups);
      for (int grp = 0; grp < p.ngroups; ++grp) {       // group images follow each other: [group][half of the pair][ky][K-step]
        const uint8_t* wsrc = (const uint8_t*)wimg + (size_t)grp * (PAIR ? 2u : 1u) * wbytes + (PAIR ? (size_t)crank * wbytes : 0);
        for (int ky = 0; ky < 3; ++ky)
          bulk_g2s(w_base + grp * wbytes + ky * nks * WSTEP, wsrc + (size_t)ky * nks * WSTEP, (uint32_t)nks * WSTEP, w_bar);
      }
#ifdef SELFC_TC_TIMING
      const long long t_pdl0 = clock64();
#endif
      pdl_wait();      // weights are static; activations come from the previous kernel in the stream
      int s = 0;
      uint32_t ph = 0;
      long long w_empty = 0;
      const long long t_start = clock64();
#ifdef SELFC_TC_TIMING
      if (p.dbg) {
        atomicMax((unsigned long long*)&p.dbg[10], (unsigned long long)(t_start - t_pdl0));
        atomicAdd((unsigned long long*)&p.dbg[11], (unsigned long long)(t_start - t_pdl0));
      }
#endif
      for (int step = rank; step < nsteps; step += nwalk) {
        const int tile = PAIR ? 2 * step + (int)crank : step;
        int tx, ty, n;
        decode_tile(p, tile, tx, ty, n);                   // n == N for the odd tile out of a pair: the box is zero-filled
        const int x0 = tx * VALID_W - 1, y0 = ty * ROWS - 1;
        for (int grp = 0; grp < p.ngroups; ++grp)
        for (int c0 = 0; c0 < nks; c0 += KPS) {
          timed_wait(empty_bar(s), ph ^ 1u, p.err, 31, w_empty);
          const int xc = P2 ? tx * (VALID_W / 2) : x0;        // P2: pair index of the halo origin (pairs start at odd x)
          if (PAIR) {
            // both CTAs' boxes complete on the LEADER's barrier, which expects the bytes of both
            if (crank == 0) mbar_expect_tx(full_bar(s), 2u * (uint32_t)STAGE);
            tma_load_5d_pair(a_base + s * STAGE, tmap, mapa_u32(full_bar(s), 0), 0, xc, y0, n, c0);
          } else {
            mbar_expect_tx(full_bar(s), (uint32_t)STAGE);
            tma_load_5d(a_base + s * STAGE, tmap, full_bar(s), 0, xc, y0, n, c0);
          }
          if (++s == NST) { s = 0; ph ^= 1u; }
        }
      }
      if (p.dbg && blockIdx.x == 0) { p.dbg[0] = w_empty; p.dbg[1] = clock64() - t_start; }
    }
  } else if (warp == 1) {
    if (PAIR && crank != 0) {
      // the peer issues no MMAs; it only tells the leader that its half of the weights is in place
      mbar_wait(w_bar, 0, p.err, 32);
      if (lane == 0) mbar_arrive_cluster(mapa_u32(wpeer_bar, 0));
    } else {
      // ===================== MMA issuer: whole warp runs the loop, one elected lane issues =====================
      constexpr uint32_t idesc = umma_idesc_bf16(PAIR ? 256 : 128, NB);
      mbar_wait(w_bar, 0, p.err, 32);
      if (PAIR) mbar_wait(wpeer_bar, 0, p.err, 36);
      int s = 0;
      uint32_t ph = 0;
      int it = 0;
      // activations: SWIZZLE_32B rows of one position (8-row atoms of 256 bytes) / P2: SWIZZLE_64B rows of a position pair (512)
      const uint32_t hi_a = (P2 || X2) ? desc_hi(512, 4) : desc_hi(256, 6);
      const uint32_t hi_b = desc_hi(128, 0);            // weights: no-swizzle core matrices, 8-row groups 128 bytes apart
      const uint32_t b_ky = (uint32_t)nks * (WSTEP >> 4);
      long long w_full = 0, w_tempty = 0;
      const long long t_start = clock64();
      for (int step = rank; step < nsteps; step += nwalk)
      for (int grp = 0; grp < p.ngroups; ++grp, ++it) {
        const int acc = it & 1;
        const uint32_t use = (uint32_t)(it >> 1);
        const uint32_t w_grp = w_base + (uint32_t)grp * 3u * (uint32_t)nks * WSTEP;
        timed_wait(tempty_bar(acc), (use & 1u) ^ 1u, p.err, 33, w_tempty);
        tc_fence_after();
        for (int c0 = 0; c0 < nks; c0 += KPS) {
          const int nk = nks - c0 < KPS ? nks - c0 : KPS;
          timed_wait(full_bar(s), ph, p.err, 34, w_full);
          tc_fence_after();
          const uint32_t a_stage = a_base + s * STAGE;
          for (int ks = 0; ks < nk; ++ks) {
            const uint32_t a_lo = desc_lo(a_stage + (uint32_t)ks * SUBB, 16);
            // B (weights): (ky, K-step) tiles; the two 8-element K core matrices are NBH/8 row-groups apart
            const uint32_t b_lo = desc_lo(w_grp + (uint32_t)(c0 + ks) * WSTEP, (NBH / 8) * 128);
            // consecutive MMAs alternate between the two M-blocks' accumulators: back-to-back accumulation into ONE
            // accumulator is a dependent chain (measured ~145 cycles per small MMA), independent accumulators pipeline
            if constexpr (X2) {
              // rows of 64 bytes: M-block mb starts 128 rows in, the ky tap one tile row (32 rows) further; the lo half of a row
              // is its second K = 16 operand (+32 bytes); the lo weight tile follows the hi tile
#pragma unroll
              for (int ky = 0; ky < 3; ++ky) {
#pragma unroll
                for (int term = 0; term < 3; ++term) {
#pragma unroll
                  for (int mb = 0; mb < MBLK; ++mb) {
                    const uint32_t d = tmem_base + (uint32_t)((acc * MBLK + mb) * ACC_STRIDE);
                    const uint64_t ad = desc_join(a_lo + (uint32_t)(((mb * 128 + ky * WT) * 64 + (term == 2 ? 32 : 0)) >> 4), hi_a);
                    const uint64_t bd = desc_join(b_lo + (uint32_t)ky * b_ky + (term == 1 ? (uint32_t)(WT_BYTES >> 4) : 0u), hi_b);
                    umma2_bf16_elect(d, ad, bd, idesc, ((c0 + ks) > 0 || ky > 0 || term > 0) ? 1u : 0u);
                  }
                }
              }
              continue;
            }
#pragma unroll
            for (int ky = 0; ky < 3; ++ky) {
#pragma unroll
              for (int mb = 0; mb < MBLK; ++mb) {
                const uint32_t d = tmem_base + (uint32_t)((acc * MBLK + mb) * ACC_STRIDE);
                // A rows = flattened halo positions: M-block mb starts at tile row 4*mb, the ky tap one tile row further
                // (P2: rows = position pairs, `mb` = which position of the pair (+32 bytes), the ky tap 16 pairs further)
                const uint64_t ad = desc_join(a_lo + (uint32_t)(P2 ? (mb * 32 + ky * (WT / 2) * 64) >> 4 : (mb * 128 + ky * WT) * 2), hi_a);
                const uint64_t bd = desc_join(b_lo + (uint32_t)ky * b_ky, hi_b);
                if (PAIR) umma2_bf16_elect(d, ad, bd, idesc, ((c0 + ks) > 0 || ky > 0) ? 1u : 0u);
                else umma_bf16_elect(d, ad, bd, idesc, ((c0 + ks) > 0 || ky > 0) ? 1u : 0u);
              }
            }
          }
          if (PAIR) umma2_commit_elect(empty_bar(s));
          else umma_commit_elect(empty_bar(s));
          if (++s == NST) { s = 0; ph ^= 1u; }
        }
        if (PAIR) umma2_commit_elect(tfull_bar(acc));
        else umma_commit_elect(tfull_bar(acc));
      }
      if (p.dbg && blockIdx.x == 0 && lane == 0) { p.dbg[2] = w_full; p.dbg[3] = w_tempty; p.dbg[4] = clock64() - t_start; }
    }
  } else {
    // ===================== epilogue warps 2..5 =====================
    const int q = warp & 3;               // TMEM lane quarter = tile row within the M-block; lane = x within the tile
    pdl_wait();
    float bias[NOUT];
#pragma unroll
    for (int j = 0; j < NOUT; ++j) bias[j] = __ldg(biasp + j);
    const size_t slab_elems = (size_t)p.slabM * 16;
    const uint32_t tempty_leader0 = PAIR ? mapa_u32(tempty_bar(0), 0) : 0u;
    int it = 0;
    long long w_tfull = 0;
    const long long t_start = clock64();
    for (int step = rank; step < nsteps; step += nwalk)
    for (int grp = 0; grp < p.ngroups; ++grp, ++it) {
      const int acc = it & 1;
      const uint32_t use = (uint32_t)(it >> 1);
      const int tile = PAIR ? 2 * step + (int)crank : step;
      const bool tile_ok = tile < p.ntiles;
      int tx, ty, n;
      decode_tile(p, tile, tx, ty, n);
      const int x = tx * VALID_W + lane;
      timed_wait(tfull_bar(acc), use & 1u, p.err, 35, w_tfull);
      tc_fence_after();
      if constexpr (P2) {
        // lane quarter q holds pairs g = 32q + lane: tile row r = g / 16, pair c = g % 16 <-> halo columns 2c, 2c+1 (global x =
        // 30 tx - 1 + column).  Output column o uses halo columns o, o+1, o+2 (kx = 0, 1, 2): this thread produces o = 2c, 2c+1.
        const int g = q * 32 + lane;
        const int r = g >> 4, c = g & 15;
        const int y = ty * ROWS + r;
        const int xo = tx * VALID_W + 2 * c;                 // global x of output 2c
        const bool row_ok = tile_ok && y < p.h && c < VALID_W / 2;
        const bool ok0 = row_ok && xo < p.w, ok1 = row_ok && xo + 1 < p.w;
        const uint32_t trow0 = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)((acc * MBLK + 0) * ACC_STRIDE);
        const uint32_t trow1 = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)((acc * MBLK + 1) * ACC_STRIDE);
        __nv_bfloat16* o = obuf + (size_t)p.out_slab * slab_elems + ((size_t)((size_t)(ok0 ? n : 0) * p.h + (ok0 ? y : 0)) * p.w + (ok0 ? xo : 0)) * 16;
        const bool drop_left = xo == 0;                      // kx = 0 term of x = 0 reads x = -1 (not a padding zero in this layout)
        const bool drop_right0 = xo == p.w - 1, drop_right1 = xo + 1 == p.w - 1;      // kx = 2 term of x = W-1 reads x = W
#pragma unroll
        for (int n0 = 0; n0 < NOUT; n0 += 16) {
          uint32_t a0[3][16], a1[3][16];
#pragma unroll
          for (int kx = 0; kx < 3; ++kx) {
            tmem_ld16(trow0 + (uint32_t)(kx * NOUT + n0), a0[kx]);
            tmem_ld16(trow1 + (uint32_t)(kx * NOUT + n0), a1[kx]);
          }
          tmem_ld_wait();
          if (n0 == NOUT - 16) {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
              if (PAIR) mbar_arrive_cluster(tempty_leader0 + 8u * (uint32_t)acc);
              else mbar_arrive(tempty_bar(acc));
            }
          }
          float v0[16], v1[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            // neighbours in the next pair (lane + 1; c = 15 wraps into the next row only for columns that are not outputs)
            const float nx0_k2 = __shfl_down_sync(0xffffffffu, __uint_as_float(a0[2][j]), 1);
            const float nx0_k1 = __shfl_down_sync(0xffffffffu, __uint_as_float(a0[1][j]), 1);
            const float nx1_k2 = __shfl_down_sync(0xffffffffu, __uint_as_float(a1[2][j]), 1);
            const float t0 = drop_left ? 0.f : __uint_as_float(a0[0][j]);
            v0[j] = lrelu02(t0 + __uint_as_float(a1[1][j]) + (drop_right0 ? 0.f : nx0_k2) + bias[n0 + j]);
            v1[j] = lrelu02(__uint_as_float(a1[0][j]) + nx0_k1 + (drop_right1 ? 0.f : nx1_k2) + bias[n0 + j]);
          }
          __nv_bfloat16* os = o + (size_t)(n0 / 16) * slab_elems;
          if (ok0) {
            uint4 lo, hi;
            lo.x = pack_bf2(v0[0], v0[1]); lo.y = pack_bf2(v0[2], v0[3]); lo.z = pack_bf2(v0[4], v0[5]); lo.w = pack_bf2(v0[6], v0[7]);
            hi.x = pack_bf2(v0[8], v0[9]); hi.y = pack_bf2(v0[10], v0[11]); hi.z = pack_bf2(v0[12], v0[13]); hi.w = pack_bf2(v0[14], v0[15]);
            *reinterpret_cast<uint4*>(os) = lo;
            *reinterpret_cast<uint4*>(os + 8) = hi;
          }
          if (ok1) {
            uint4 lo, hi;
            lo.x = pack_bf2(v1[0], v1[1]); lo.y = pack_bf2(v1[2], v1[3]); lo.z = pack_bf2(v1[4], v1[5]); lo.w = pack_bf2(v1[6], v1[7]);
            hi.x = pack_bf2(v1[8], v1[9]); hi.y = pack_bf2(v1[10], v1[11]); hi.z = pack_bf2(v1[12], v1[13]); hi.w = pack_bf2(v1[14], v1[15]);
            *reinterpret_cast<uint4*>(os + 16) = lo;
            *reinterpret_cast<uint4*>(os + 24) = hi;
          }
        }
      } else {
#pragma unroll
        for (int mb = 0; mb < MBLK; ++mb) {
          const int y = ty * ROWS + mb * 4 + q;
          const bool ok = tile_ok && lane < VALID_W && x < p.w && y < p.h;
          const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)((acc * MBLK + mb) * ACC_STRIDE);
          constexpr int ROWE = X2 ? 32 : 16;       // bf16 elements per (position, slab) row
          __nv_bfloat16* o = obuf + (size_t)p.out_slab * slab_elems * (ROWE / 16) +
                             ((size_t)((size_t)(ok ? n : 0) * p.h + (ok ? y : 0)) * p.w + (ok ? x : 0)) * ROWE;
  #pragma unroll
          for (int n0 = 0; n0 < NOUT; n0 += 16) {
            uint32_t r0[16], r1[16], r2[16];
            tmem_ld16(trow + (uint32_t)n0, r0);
            tmem_ld16(trow + (uint32_t)(NOUT + n0), r1);
            tmem_ld16(trow + (uint32_t)(2 * NOUT + n0), r2);
            tmem_ld_wait();
            if (mb == MBLK - 1 && n0 == NOUT - 16) {
              // both accumulators of this tile are in registers: the next-but-one tile's MMAs may start
              tc_fence_before();
              __syncwarp();
              if (lane == 0) {
                if (PAIR) mbar_arrive_cluster(tempty_leader0 + 8u * (uint32_t)acc);
                else mbar_arrive(tempty_bar(acc));
              }
            }
            float v[16];
  #pragma unroll
            for (int j = 0; j < 16; ++j) {
              const float a1 = __shfl_down_sync(0xffffffffu, __uint_as_float(r1[j]), 1);
              const float a2 = __shfl_down_sync(0xffffffffu, __uint_as_float(r2[j]), 2);
              v[j] = __uint_as_float(r0[j]) + a1 + a2;
              if (!X2 || p.accF == nullptr) v[j] = lrelu02(v[j] + bias[n0 + j]);
            }
            if (X2 && p.accF != nullptr) {
              if (ok) {
                float* of = p.accF + dense_off((long long)((size_t)((size_t)n * p.h + y) * p.w + x), p.acc_off + 32 * grp + n0, p.acc_pitch, p.acc_slabM);
  #pragma unroll
                for (int j = 0; j < 16; j += 4) {
                  if (32 * grp + n0 + j < p.acc_n) {
                    float4 t = *reinterpret_cast<const float4*>(of + j);
                    t.x += v[j]; t.y += v[j + 1]; t.z += v[j + 2]; t.w += v[j + 3];
                    *reinterpret_cast<float4*>(of + j) = t;
                  }
                }
              }
            } else if (ok) {
              if constexpr (X2) {
                __nv_bfloat16* os = o + (size_t)(n0 / 16) * slab_elems * 2;
                x2_store8(os, v);              // channels 0..7: hi at +0, lo at +16 elements
                x2_store8(os + 8, v + 8);      // channels 8..15
              } else {
                uint4 lo, hi;
                lo.x = pack_bf2(v[0], v[1]); lo.y = pack_bf2(v[2], v[3]); lo.z = pack_bf2(v[4], v[5]); lo.w = pack_bf2(v[6], v[7]);
                hi.x = pack_bf2(v[8], v[9]); hi.y = pack_bf2(v[10], v[11]); hi.z = pack_bf2(v[12], v[13]); hi.w = pack_bf2(v[14], v[15]);
                __nv_bfloat16* os = o + (size_t)(n0 / 16) * slab_elems;
                *reinterpret_cast<uint4*>(os) = lo;
                *reinterpret_cast<uint4*>(os + 8) = hi;
              }
            }
          }
        }
      }
    }
    if (p.dbg && blockIdx.x == 0 && threadIdx.x == 64) { p.dbg[6] = w_tfull; p.dbg[8] = clock64() - t_start; p.dbg[9] = it; }
#ifdef SELFC_TC_TIMING
    if (p.dbg && threadIdx.x == 64) {      // busy time of every CTA's epilogue (first accumulator ready -> last store), max / sum
      const long long busy = clock64() - t_start - w_tfull;
      atomicMax((unsigned long long*)&p.dbg[12], (unsigned long long)busy);
      atomicAdd((unsigned long long*)&p.dbg[13], (unsigned long long)busy);
      atomicMax((unsigned long long*)&p.dbg[14], (unsigned long long)(clock64() - t_start));
      atomicAdd((unsigned long long*)&p.dbg[15], (unsigned long long)(clock64() - t_start));
    }
#endif
  }

  tc_fence_before();
  __syncthreads();
  if (PAIR) cluster_sync_all();      // neither CTA leaves (or frees tensor memory) while the other may still signal it
  if (warp == 1) {
    tc_fence_after();
    if (PAIR) tmem_dealloc2(tmem_base, (uint32_t)TMEM_COLS);
    else tmem_dealloc(tmem_base, (uint32_t)TMEM_COLS);
  }
}

// wref [32][cin_ref][3][3] fp32 -> bf16 B-operand image [ky][kstep][kcore(2)][ngroup(12)][r%8][k%8], row r = kx*32 + n;
// img_pair: the same rows split between the CTAs of a pair, [half(2)][ky][kstep][kcore(2)][ngroup(6)][r%8][k%8], half = r / 48
// img_x2 (BF16X3 mode): the pair image with a hi and a lo tile per (ky, K-step), [half(2)][ky][kstep][hi|lo][kcore(2)][ngroup(6)][r%8][k%8],
// hi = bf16(w), lo = bf16(w - hi)
struct PackTc3Job {
  const float* wref;
  const float* bref;       // conv bias [32] -> bias (the epilogue's copy)
  __nv_bfloat16* img;
  __nv_bfloat16* img_pair;
  __nv_bfloat16* img_x2;
  float* bias;
  int cin_ref, cin_buf, xreal, xpad;
};

__device__ __forceinline__ void pack_tc3_body(const PackTc3Job& j, int idx) {
  const float* __restrict__ wref = j.wref;
  __nv_bfloat16* __restrict__ img = j.img;
  __nv_bfloat16* __restrict__ img_pair = j.img_pair;
  __nv_bfloat16* __restrict__ img_x2 = j.img_x2;
  const int cin_ref = j.cin_ref, cin_buf = j.cin_buf, xreal = j.xreal, xpad = j.xpad;
  const int nks = cin_buf / 16;
  const int total = 3 * cin_buf * NB;
  if (idx < NOUT) j.bias[idx] = j.bref[idx];
  if (idx >= total) return;
  const int r = idx % NB;
  const int c = (idx / NB) % cin_buf;
  const int ky = idx / (NB * cin_buf);
  const int kx = r / 32, n = r % 32;
  const int cref = c < xreal ? c : (c < xpad ? -1 : c - xpad + xreal);
  float v = 0.f;
  if (cref >= 0 && cref < cin_ref) v = wref[((size_t)n * cin_ref + cref) * 9 + ky * 3 + kx];
  const int ks = c / 16, kk = c % 16;
  const size_t off = (size_t)(ky * nks + ks) * (WTILE_BYTES / 2) + (size_t)((kk / 8) * (NB / 8) + r / 8) * 64 + (r % 8) * 8 + (kk % 8);
  img[off] = __float2bfloat16_rn(v);
  const int half = r / (NB / 2), rh = r % (NB / 2);
  const size_t half_elems = (size_t)3 * nks * (WTILE_BYTES / 4);
  const size_t offp = half * half_elems + (size_t)(ky * nks + ks) * (WTILE_BYTES / 4) + (size_t)((kk / 8) * (NB / 16) + rh / 8) * 64 +
                      (rh % 8) * 8 + (kk % 8);
  img_pair[offp] = __float2bfloat16_rn(v);
  if (img_x2 != nullptr) {
    const size_t tile = (size_t)(WTILE_BYTES / 4);
    const size_t inner = (size_t)((kk / 8) * (NB / 16) + rh / 8) * 64 + (rh % 8) * 8 + (kk % 8);
    const size_t offx = half * (2 * half_elems) + (size_t)((ky * nks + ks) * 2) * tile + inner;
    __nv_bfloat16 hi, lo;
    x2_split(v, hi, lo);
    img_x2[offx] = hi;
    img_x2[offx + tile] = lo;
  }
}

__global__ void pack_tc3_kernel(const PackTc3Job j) { pack_tc3_body(j, blockIdx.x * blockDim.x + threadIdx.x); }

// every recorded job in one launch (pack_batch.h)
__global__ void pack_tc3_multi_kernel(const PackTc3Job* __restrict__ jobs, const int* __restrict__ first, int njobs) {
  const int ji = pack_find_job(first, njobs, blockIdx.x);
  const PackTc3Job j = jobs[ji];
  pack_tc3_body(j, (blockIdx.x - __ldg(first + ji)) * blockDim.x + threadIdx.x);
}

// Input-gradient images of a dense block (training, BF16X3 mode), one per 32-channel SLOT of its buffer.  The gradient w.r.t. the
// buffer channels c of a slot is the sum over the later convs k of conv_k^T(g_k): gx[c] = sum_k sum_{n,ky,kx} Wf_k[(2-ky)*3 + (2-kx)][c][n]
// . g_k[n] at (y + ky - 1, x + kx - 1), Wf_k = conv k's forward weights in buffer-channel order [tap][cin_k][32] (conv_simt pack).  With the
// output gradients of the contributing convs CONCATENATED in K -- [g_4 | g_3 | ...], 32 channels each, the order the backward produces
// them in -- that is ONE conv with K = 32 * nconv per slot, so every gradient channel is written once instead of once per later conv.
// Image j of a slot holds rows r = kx * 32 + (c - c0 - 32 j) in the pair layout of pack_tc3_kernel's img_x2 with 2 * nconv K-steps;
// channels beyond the slot are zero rows.
struct DgradSlotSrc {
  const float* wf[4];      // forward packs of conv1..conv4
  int cin[4];              // their buffer-channel counts
};
struct PackSlotJob {
  DgradSlotSrc src;
  __nv_bfloat16* img;
  int c0, ncover, nconv, ngroups;
};
__device__ __forceinline__ void pack_tc3_dgrad_slot_body(const PackSlotJob& job, int idx) {
  const DgradSlotSrc& src = job.src;
  __nv_bfloat16* __restrict__ img = job.img;
  const int c0 = job.c0, ncover = job.ncover, nconv = job.nconv, ngroups = job.ngroups;
  const int K = 32 * nconv;
  const int per = 3 * K * NB;                        // (ky, kk, r) of one group
  if (idx >= per * ngroups) return;
  const int j = idx / per, e = idx - j * per;
  const int r = e % NB;
  const int kk = (e / NB) % K;                       // K index: conv 4 - kk / 32 (conv4's gradient first), its output channel kk % 32
  const int ky = e / (NB * K);
  const int kc = 3 - kk / 32, n = kk % 32;
  const int kx = r / 32, cl = 32 * j + r % 32;
  float v = 0.f;
  if (cl < ncover) v = src.wf[kc][((size_t)((2 - ky) * 3 + (2 - kx)) * src.cin[kc] + c0 + cl) * 32 + n];
  const int nks = 2 * nconv;
  const int ks = kk / 16, k16 = kk % 16;
  const int half = r / (NB / 2), rh = r % (NB / 2);
  const size_t tile = (size_t)(WTILE_BYTES / 4);
  const size_t half_elems = (size_t)3 * nks * tile;
  const size_t inner = (size_t)((k16 / 8) * (NB / 16) + rh / 8) * 64 + (rh % 8) * 8 + (k16 % 8);
  const size_t off = (size_t)j * (4 * half_elems) + half * (2 * half_elems) + (size_t)((ky * nks + ks) * 2) * tile + inner;
  __nv_bfloat16 hi, lo;
  x2_split(v, hi, lo);
  img[off] = hi;
  img[off + tile] = lo;
}

__global__ void pack_tc3_dgrad_slot_kernel(const PackSlotJob j) { pack_tc3_dgrad_slot_body(j, blockIdx.x * blockDim.x + threadIdx.x); }

// every recorded job in one launch (pack_batch.h)
__global__ void pack_tc3_dgrad_slot_multi_kernel(const PackSlotJob* __restrict__ jobs, const int* __restrict__ first, int njobs) {
  const int ji = pack_find_job(first, njobs, blockIdx.x);
  const PackSlotJob j = jobs[ji];
  pack_tc3_dgrad_slot_body(j, (blockIdx.x - __ldg(first + ji)) * blockDim.x + threadIdx.x);
}

}  // namespace tc3

size_t tc3_dgrad_slot_image_bytes(int nconv) { return (size_t)2 * 3 * (2 * nconv) * tc3::WTILE_BYTES; }      // per 32-channel group: hi + lo, 3 ky

int pack_tc3_dgrad_slot_images(const float* const wf[4], const int cin[4], void* img, int c0, int ncover, int nconv, cudaStream_t st) {
  SELFC_CHECK_ARG(nconv >= 1 && nconv <= 4 && ncover >= 1 && c0 >= 0, "dgrad slot images: bad slot");
  tc3::PackSlotJob j;
  memset(&j, 0, sizeof(j));      // (padding bytes too: the job tables are compared bytewise)
  for (int k = 0; k < 4; ++k) { j.src.wf[k] = wf[k]; j.src.cin[k] = cin[k]; }
  const int ngroups = cdiv(ncover, 32);
  const int total = 3 * 32 * nconv * tc3::NB * ngroups;
  j.img = reinterpret_cast<__nv_bfloat16*>(img); j.c0 = c0; j.ncover = ncover; j.nconv = nconv; j.ngroups = ngroups;
  if (PackBatch* pb = pack_batch_current()) {
    pb->slot[pb->point].add(j, (int)cdiv(total, 256));
    return 0;
  }
  tc3::pack_tc3_dgrad_slot_kernel<<<cdiv(total, 256), 256, 0, st>>>(j);
  SELFC_LAUNCH_CHECK("pack_tc3_dgrad_slot_kernel");
  return 0;
}

int flush_pack_dgrad_slot(JobTable& t, cudaStream_t st) {
  if (t.njobs() <= 0) return 0;
  const void* jobs = nullptr;
  const int* first = nullptr;
  SELFC_CUDA(t.sync(st, &jobs, &first));
  tc3::pack_tc3_dgrad_slot_multi_kernel<<<t.first.back(), 256, 0, st>>>(static_cast<const tc3::PackSlotJob*>(jobs), first, t.njobs());
  SELFC_LAUNCH_CHECK("pack_tc3_dgrad_slot_multi_kernel");
  return 0;
}

// ---- host-side helpers shared by the tcgen05 kernels -----------------------------------------------------------
namespace tc {
EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

static int* g_err_flag[64] = {};

int* err_flag_for_device() {
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64) return nullptr;
  if (!g_err_flag[dev]) {
    if (cudaMalloc(&g_err_flag[dev], sizeof(int)) != cudaSuccess) return nullptr;
    cudaMemset(g_err_flag[dev], 0, sizeof(int));
  }
  return g_err_flag[dev];
}

static long long* g_dbg_dev = nullptr;
static long long g_dbg_tags[4096];
static int g_dbg_n = 0;
bool debug_slots() {
  static int on = -1;
  if (on < 0) {
    const char* e = getenv("SELFC_TC_DBG");
    on = (e && atoi(e) != 0) ? 1 : 0;
  }
  return on == 1;
}
long long* debug_next_slot(long long tag) {
  if (!g_dbg_dev) {
    if (cudaMalloc(&g_dbg_dev, 4096 * 16 * sizeof(long long)) != cudaSuccess) return nullptr;
    cudaMemset(g_dbg_dev, 0, 4096 * 16 * sizeof(long long));
  }
  if (g_dbg_n >= 4096) return nullptr;
  g_dbg_tags[g_dbg_n] = tag;
  return g_dbg_dev + 16 * (g_dbg_n++);
}
int debug_read(long long* out, int cap) {
  cudaDeviceSynchronize();
  int n = g_dbg_n < cap ? g_dbg_n : cap;
  for (int i = 0; i < n; ++i) {
    out[17 * i] = g_dbg_tags[i];
    cudaMemcpy(out + 17 * i + 1, g_dbg_dev + 16 * i, 16 * sizeof(long long), cudaMemcpyDeviceToHost);
  }
  g_dbg_n = 0;
  if (g_dbg_dev) cudaMemset(g_dbg_dev, 0, 4096 * 16 * sizeof(long long));
  return n;
}

// sweep direction of the next tcgen05 conv launch: alternates per launch (SELFC_ZIGZAG=0: always forward)
int next_direction() {
  static int on = -1;
  static int dir = 0;
  if (on < 0) {
    const char* e = getenv("SELFC_ZIGZAG");
    on = (e && atoi(e) == 0) ? 0 : 1;
  }
  if (!on) return 0;
  dir ^= 1;
  return dir;
}

bool pdl_enabled() {
  static int on = -1;
  if (on < 0) {
    const char* e = getenv("SELFC_NO_PDL");
    on = (e && atoi(e) != 0) ? 0 : 1;
  }
  return on == 1;
}

int num_sms() {
  static int n = 0;
  if (!n) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
  }
  return n;
}
}  // namespace tc

int pack_tc_weights(TcConvW& w, const float* wref, const float* bref, int cin_ref, int cin_buf, int xreal, int xpad, cudaStream_t st,
                    bool x2) {
  SELFC_CHECK_ARG(cin_buf % 16 == 0 && cin_buf <= tc3::MAX_CIN, "conv3x3_tc: cin %d must be a multiple of 16 and <= %d", cin_buf,
                  tc3::MAX_CIN);
  const size_t bytes = (size_t)3 * (cin_buf / 16) * tc3::WTILE_BYTES;
  if (w.img == nullptr || w.img_bytes != bytes) {
    free_tc_weights(w);
    SELFC_CUDA(cudaMalloc(&w.img, bytes));
    SELFC_CUDA(cudaMalloc(&w.img_pair, bytes));
    SELFC_CUDA(cudaMalloc(&w.bias, tc3::NOUT * sizeof(float)));
    w.img_bytes = bytes;
  }
  if (x2 && w.img_x2 == nullptr) SELFC_CUDA(cudaMalloc(&w.img_x2, 2 * bytes));
  w.cin_buf = cin_buf;
  const int total = 3 * cin_buf * tc3::NB;
  const tc3::PackTc3Job j{wref, bref, reinterpret_cast<__nv_bfloat16*>(w.img), reinterpret_cast<__nv_bfloat16*>(w.img_pair),
                          x2 ? reinterpret_cast<__nv_bfloat16*>(w.img_x2) : nullptr, w.bias, cin_ref, cin_buf, xreal, xpad};
  if (PackBatch* pb = pack_batch_current()) {
    pb->tc3[pb->point].add(j, (int)cdiv(total, 256));
    return 0;
  }
  tc3::pack_tc3_kernel<<<cdiv(total, 256), 256, 0, st>>>(j);
  SELFC_LAUNCH_CHECK("pack_tc3_kernel");
  return 0;
}

int flush_pack_tc3(JobTable& t, cudaStream_t st) {
  if (t.njobs() <= 0) return 0;
  const void* jobs = nullptr;
  const int* first = nullptr;
  SELFC_CUDA(t.sync(st, &jobs, &first));
  tc3::pack_tc3_multi_kernel<<<t.first.back(), 256, 0, st>>>(static_cast<const tc3::PackTc3Job*>(jobs), first, t.njobs());
  SELFC_LAUNCH_CHECK("pack_tc3_multi_kernel");
  return 0;
}

void free_tc_weights(TcConvW& w) {
  if (w.img) cudaFree(w.img);
  if (w.img_pair) cudaFree(w.img_pair);
  if (w.img_x2) cudaFree(w.img_x2);
  if (w.bias) cudaFree(w.bias);
  w.img = nullptr;
  w.img_pair = nullptr;
  w.img_x2 = nullptr;
  w.bias = nullptr;
  w.img_bytes = 0;
}

int launch_conv3x3_tc(const TcConvW& w, __nv_bfloat16* buf, long long slabM, int cin, int out_off, int N, int h, int wd, cudaStream_t st,
                      const TcConvW* w2, __nv_bfloat16* buf2, bool x2, const TcAccum* acc) {
  SELFC_CHECK_ARG(acc == nullptr || (x2 && w2 == nullptr && acc->out != nullptr && acc->pitch % 4 == 0 && acc->off % 4 == 0 && acc->n % 4 == 0 &&
                                     acc->n >= 4 && acc->ngroups >= 1 && acc->ngroups <= 6 && acc->n <= 32 * acc->ngroups && aligned16(acc->out)),
                  "conv3x3_tc: the accumulate epilogue belongs to a single (hi, lo) problem with 16-byte aligned fp32 rows");
  SELFC_CHECK_ARG(acc == nullptr || acc->slabM == 0 || (acc->off % 16 == 0 && acc->pitch % 16 == 0), "conv3x3_tc: slab-planar accumulate buffer");
  SELFC_CHECK_ARG(w.img != nullptr && cin == w.cin_buf, "conv3x3_tc: weights not packed for cin=%d", cin);
  SELFC_CHECK_ARG(!x2 || (w.img_x2 != nullptr && (w2 == nullptr || w2->img_x2 != nullptr) && ((uintptr_t)buf & 63) == 0 &&
                          (buf2 == nullptr || ((uintptr_t)buf2 & 63) == 0)),
                  "conv3x3_tc: the (hi, lo) form needs its weight image and 64-byte aligned buffers");
  SELFC_CHECK_ARG(out_off % 16 == 0 && aligned16(buf) && slabM == (long long)N * h * wd, "conv3x3_tc: slab layout / alignment");
  const bool dual = w2 != nullptr;
  SELFC_CHECK_ARG(!dual || (w2->img != nullptr && w2->cin_buf == cin && buf2 != nullptr && aligned16(buf2) && buf2 != buf),
                  "conv3x3_tc: the second problem must have the same shape and its own buffer");
  tc::EncodeTiledFn encode = tc::get_encode_fn();
  if (!encode) {
    set_error("cuTensorMapEncodeTiled is not available from this driver");
    return SELFC_E_CUDA;
  }
  static int kps_pref = 0;     // SELFC_TC3_KPS: K-steps (16-channel slabs) per pipeline stage / TMA box
  if (!kps_pref) {
    const char* e = getenv("SELFC_TC3_KPS");
    kps_pref = e ? atoi(e) : 2;
    if (kps_pref < 1 || kps_pref > 4) kps_pref = 2;
  }
  // SELFC_TC3_PAIR=0: one-CTA kernel (cta_group::1) instead of CTA pairs.  The pair kernel's cheaper MMA issue only pays together
  // with the position-pair TMA rows (SELFC_TC3_P2): each alone leaves the other limit in place (profiles/r2_conv3x3_waits.md).
  static int pair_pref = -1;
  if (pair_pref < 0) {
    const char* e = getenv("SELFC_TC3_PAIR");
    pair_pref = (e && atoi(e) == 0) ? 0 : 1;
  }
  const bool pair = x2 || (pair_pref == 1 && w.img_pair != nullptr && (!dual || w2->img_pair != nullptr));
  const int nks = cin / 16;
  int kps = kps_pref < nks ? kps_pref : nks;
  const int sub_bytes = x2 ? 2 * tc3::SUB_BYTES : tc3::SUB_BYTES;
  // both problems' weights have the same size; x2: hi + lo images, half of each per CTA
  const int ngroups = acc != nullptr ? acc->ngroups : 1;
  const int fixed = tc3::BAR_BYTES + (int)(x2 ? w.img_bytes * ngroups : (pair ? w.img_bytes / 2 : w.img_bytes)) + 1024;
  while (kps > 1 && (227 * 1024 - fixed) / (kps * sub_bytes) < 3) --kps;
  // SELFC_TC3_P2=0: one position per TMA row (SWIZZLE_32B) instead of position pairs.  Pairs need an even width.
  static int p2_pref = -1;
  if (p2_pref < 0) {
    const char* e = getenv("SELFC_TC3_P2");
    p2_pref = (e && atoi(e) == 0) ? 0 : 1;
  }
  const bool p2 = !x2 && p2_pref == 1 && (wd % 2) == 0;
  CUtensorMap tmap, tmap2;
  CUresult r;
  if (x2) {
    // rows = one position's 16 channels as [16 x hi | 16 x lo] (64 bytes)
    const cuuint64_t gdim[5] = {32, (cuuint64_t)wd, (cuuint64_t)h, (cuuint64_t)N, (cuuint64_t)nks};
    const cuuint64_t gstr[4] = {64, (cuuint64_t)wd * 64, (cuuint64_t)h * wd * 64, (cuuint64_t)slabM * 64};
    const cuuint32_t box[5] = {32, (cuuint32_t)tc3::WT, (cuuint32_t)tc3::HT, 1, (cuuint32_t)kps};
    const cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    r = encode(&tmap, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, buf, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r == CUDA_SUCCESS)
      r = encode(&tmap2, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, dual ? buf2 : buf, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                 CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  } else if (p2) {
    // rows = pairs of positions (x odd, x + 1): based one position before the buffer, W/2 + 1 pairs per image row
    const cuuint64_t gdim[5] = {32, (cuuint64_t)wd / 2 + 1, (cuuint64_t)h, (cuuint64_t)N, (cuuint64_t)nks};
    const cuuint64_t gstr[4] = {64, (cuuint64_t)wd * 32, (cuuint64_t)h * wd * 32, (cuuint64_t)slabM * 32};
    const cuuint32_t box[5] = {32, (cuuint32_t)tc3::WT / 2, (cuuint32_t)tc3::HT, 1, (cuuint32_t)kps};
    const cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    r = encode(&tmap, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, buf - 16, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r == CUDA_SUCCESS)
      r = encode(&tmap2, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, (dual ? buf2 : buf) - 16, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                 CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  } else {
    const cuuint64_t gdim[5] = {16, (cuuint64_t)wd, (cuuint64_t)h, (cuuint64_t)N, (cuuint64_t)nks};
    const cuuint64_t gstr[4] = {32, (cuuint64_t)wd * 32, (cuuint64_t)h * wd * 32, (cuuint64_t)slabM * 32};
    const cuuint32_t box[5] = {16, (cuuint32_t)tc3::WT, (cuuint32_t)tc3::HT, 1, (cuuint32_t)kps};
    const cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    r = encode(&tmap, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, buf, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r == CUDA_SUCCESS)
      r = encode(&tmap2, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, dual ? buf2 : buf, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                 CU_TENSOR_MAP_SWIZZLE_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  }
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed with CUresult %d (%dx%dx%d, %d slabs)", (int)r, N, h, wd, nks);
    return SELFC_E_CUDA;
  }
  tc3::Params p;
  memset(&p, 0, sizeof(p));
  p.wimg = x2 ? w.img_x2 : (pair ? w.img_pair : w.img);
  p.bias = w.bias;
  p.buf = buf;
  p.nprob = dual ? 2 : 1;
  p.wimg2 = dual ? (x2 ? w2->img_x2 : (pair ? w2->img_pair : w2->img)) : p.wimg;
  p.bias2 = dual ? w2->bias : w.bias;
  p.buf2 = dual ? buf2 : buf;
  p.slabM = slabM;
  p.out_slab = out_off / 16;
  p.N = N;
  p.h = h;
  p.w = wd;
  p.nks = nks;
  p.kps = kps;
  p.tiles_x = cdiv(wd, tc3::VALID_W);
  p.tiles_y = cdiv(h, tc3::ROWS);
  p.ntiles = p.tiles_x * p.tiles_y * N;
  p.err = tc::err_flag_for_device();
  p.ngroups = ngroups;
  if (acc != nullptr) { p.accF = acc->out; p.acc_pitch = acc->pitch; p.acc_off = acc->off; p.acc_n = acc->n; p.acc_slabM = acc->slabM; }
  if (p.ntiles == 0) return 0;
  p.rev = tc::next_direction();
  if (tc::debug_slots()) p.dbg = tc::debug_next_slot(9000000 + (dual ? 100000 : 0) + nks);
  const int stage = kps * sub_bytes;
  int nst = (227 * 1024 - fixed) / stage;
  if (nst > tc3::NSTAGE_MAX) nst = tc3::NSTAGE_MAX;
  SELFC_CHECK_ARG(nst >= 2, "conv3x3_tc: no room for the activation pipeline (cin %d)", cin);
  p.nst = nst;
  const int smem = fixed + nst * stage;
  static bool smem_set[64] = {};           // per device: function attributes belong to the device's context
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev >= 0 && dev < 64 && !smem_set[dev]) {
    SELFC_CUDA(cudaFuncSetAttribute(tc3::conv3x3_tc3_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    SELFC_CUDA(cudaFuncSetAttribute(tc3::conv3x3_tc3_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    SELFC_CUDA(cudaFuncSetAttribute(tc3::conv3x3_tc3_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    SELFC_CUDA(cudaFuncSetAttribute(tc3::conv3x3_tc3_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    SELFC_CUDA(cudaFuncSetAttribute(tc3::conv3x3_tc3_kernel<true, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    smem_set[dev] = true;
  }
  const int nsm = tc::num_sms();
  if (pair) {
    // scheduling units are CTA pairs, each taking two tiles per step; dual launches give every problem the same number of pairs
    const int nsteps = (p.ntiles + 1) / 2;
    int npairs = p.nprob * nsteps < nsm / 2 ? p.nprob * nsteps : nsm / 2;
    if (dual && (npairs & 1)) --npairs;
    if (x2)
      SELFC_CUDA(tc::launch_pdl_pairs(tc3::conv3x3_tc3_kernel<true, false, true>, 2 * npairs, tc3::THREADS, smem, st, tmap, tmap2, p));
    else
      SELFC_CUDA(tc::launch_pdl_pairs(p2 ? tc3::conv3x3_tc3_kernel<true, true> : tc3::conv3x3_tc3_kernel<true, false>, 2 * npairs,
                                      tc3::THREADS, smem, st, tmap, tmap2, p));
  } else {
    int grid = p.nprob * p.ntiles < nsm ? p.nprob * p.ntiles : nsm;
    if (dual && (grid & 1)) --grid;                   // CTA parity selects the problem: both halves get the same CTA count
    SELFC_CUDA(tc::launch_pdl(p2 ? tc3::conv3x3_tc3_kernel<false, true> : tc3::conv3x3_tc3_kernel<false, false>, grid, tc3::THREADS,
                              smem, st, tmap, tmap2, p));
  }
  SELFC_LAUNCH_CHECK("conv3x3_tc3_kernel");
  return 0;
}

}  // namespace selfc
