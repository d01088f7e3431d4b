// conv1..convL of a D2DTInput dense block (Subnet_constructor.py:102-105,126-129) as ONE launch: the growth channels
// x1..x_{L-1} never leave the SM between the layers.  BF16 mode, tcgen05 / TMEM / TMA, CTA pairs (cta_group::2).
//
// Unfused (conv_tc3.cu) every layer re-reads [X | x1 | ...] from HBM: (cin + 32(k-1)) channels in, 32 out, per layer.  Here a
// CTA owns a vertical STRIP of the image -- 128 positions wide, 120 of them outputs of the last layer -- and slides down it one
// row per step with the four layers software-pipelined one behind the other ("line buffers"):
//
//     step s issues, in this order:   conv1 row s,   conv3 row s-3,   conv2 row s-1,   conv4 row s-4          (L = 4)
//
// A row of a layer is one accumulator: D[p][kx*32+n] += sum_c A[row+ky-1][p][c] * W[ky,kx][n][c] (M = 128 positions of this
// CTA + 128 of its pair, N = 96 = the three kx taps stacked, K = 16), out[p][n] = D[p-1][kx=0] + D[p][kx=1] + D[p+1][kx=2]
// -- the same arithmetic, in the same order, as conv3x3_tc3_kernel, so the results are bit-identical to the unfused path.
// The A operands: X rows come from HBM by TMA (SWIZZLE_32B, one box per row: 128 positions x nx slabs) into an 8-row ring in
// shared memory; the growth rows x1..x3 are written by the epilogue warps as bf16 straight into TENSOR MEMORY (tcgen05.st) and
// read from there by the MMAs (A-in-TMEM form: lane = position, 8 columns per 16 channels; issue rate 48 cycles for N = 96,
// profiles/r2b_ubench_mma_tmem_a.txt) -- rings of 6 / 5 / 3 rows = 224 columns next to three 96-column accumulators.
// No vertical recompute (only 2 x (L-1) rows where a CTA's row range starts / ends), horizontal recompute 8 of 128 positions
// (the same 15/16 efficiency as the 30-of-32 tiles of the unfused kernel).  Out-of-image rows are skipped MMAs, out-of-image
// columns are TMA zero-fill (X) or zeros written by the epilogue (growth rows): exactly the zero padding of the reference.
// Every layer's 32 outputs are also stored to the dense buffer in HBM once (conv5 reads them): per pixel a block moves
// (cin + 128) bf16 channels instead of (4 cin + 192 + 128).
//
// The order above puts an independent row between a producer and its consumer (conv2 row s-1 needs x1 row s; conv4 row s-4
// needs x3 row s-3), which hides the epilogue latency; ring sizes follow from it (a slot is overwritten only after every MMA
// issued before the overwriting row's own MMAs has completed -- tcgen05.commit semantics).
//
// Per CTA (320 threads): warp 0 TMA producer, warp 1 MMA issuer (leader CTA only), warps 2..9 epilogue -- two warpgroups, each
// taking 16 of a row's 32 output channels (halves the producer->consumer latency).  The lane +-1 neighbours of the kx
// combination are warp shuffles inside a 32-position quarter and a 512-byte shared-memory exchange across quarters.
#include <cuda.h>
#include <stdlib.h>
#include <string.h>
#include <type_traits>

#include "common.cuh"
#include "conv_tc.h"
#include "tc_ptx.cuh"

namespace selfc {
namespace dbf {

using namespace tc;

constexpr int MPOS = 128;                     // positions per strip row = M of one CTA's half of the pair MMA
constexpr int HALO = 4;                       // strip buffer index 0 <-> x = x0 - HALO
constexpr int VALID = MPOS - 2 * HALO;        // 120 outputs of the last layer per strip row
constexpr int NOUT = 32;
constexpr int NB = 3 * NOUT;                  // 96
constexpr int NBH = NB / 2;                   // weight rows per CTA of the pair
constexpr int WT_BYTES = NBH * 16 * 2;        // one (ky, K-step) B tile of one CTA: 1536
constexpr int SLAB_ROW = MPOS * 32;           // one 16-channel slab of one strip row: 4 KB
constexpr int NXR = 8;                        // X ring slots (rows)
constexpr int NACC = 3;                       // accumulators in rotation
constexpr int MAXL = 4;
constexpr int THREADS = 320;
constexpr int TMEM_COLS = 512;
constexpr int RING_COL0 = NACC * NB;          // 288
constexpr int BAR_BYTES = 512;
constexpr int NRING_MAX = 14;                 // growth-ring slots of all layers (6 + 5 + 3 for L = 4)
constexpr int XBUF_BYTES = 2 * 2 * 4 * 2 * 16 * 4;      // [warpgroup][parity][quarter][side][16] floats
constexpr int BIAS_BYTES = MAXL * NOUT * 4;

// schedule tables (see the header comment); L = number of fused layers
__host__ __device__ constexpr int lag_of(int L, int j) { return L == 4 ? (j == 0 ? 0 : j == 1 ? 1 : j == 2 ? 3 : 4) : (j == 0 ? 0 : j == 1 ? 1 : 3); }
__host__ __device__ constexpr int order_of(int L, int oi) { return L == 4 ? (oi == 0 ? 0 : oi == 1 ? 2 : oi == 2 ? 1 : 3) : (oi == 0 ? 0 : oi == 1 ? 2 : 1); }
__host__ __device__ constexpr int ring_of(int L, int j) { return L == 4 ? (j == 0 ? 6 : j == 1 ? 5 : 3) : (j == 0 ? 5 : 3); }
__host__ __device__ constexpr int ringcol_of(int L, int j) {
  return RING_COL0 + (j == 0 ? 0 : j == 1 ? 16 * ring_of(L, 0) : 16 * (ring_of(L, 0) + ring_of(L, 1)));
}
__host__ __device__ constexpr int ringslot0_of(int L, int j) { return (ringcol_of(L, j) - RING_COL0) / 16; }      // first global slot index
static_assert(ringcol_of(4, 2) + 16 * ring_of(4, 2) <= TMEM_COLS, "TMEM budget");
static_assert(ringslot0_of(4, 2) + ring_of(4, 2) <= 14, "ring slot barriers");

struct Params {
  const void* wimg[2][MAXL];     // per problem, per layer: TcConvW::img_pair (two halves of the B image)
  const float* bias[2][MAXL];
  __nv_bfloat16* buf[2];
  int nprob, nx;                 // nx = 16-channel slabs of X
  long long slabM;
  int N, h, w, S, ncol;          // S strips per image row, ncol = N * S strip columns
  int piece_len, total_pr;       // a CTA pair owns piece_len consecutive rows of the (column pair, row) sequence
  int* err;
};

__device__ __forceinline__ void tma_load_4d_pair(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
// A operand in tensor memory (lane = row, 8 columns = 16 bf16 of K)
__device__ __forceinline__ void umma2_ts_bf16_elect(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p, q;\n\t"
      "elect.sync _|q, 0xffffffff;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "@q tcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t r[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(r[0]), "r"(r[1]),
               "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory"); }
__device__ __forceinline__ uint32_t pack_bf2(float a, float b) {
  __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&v);
}

template <int V> using IC = std::integral_constant<int, V>;

// A piece = rows [pr_begin, pr_end) of the flattened (column pair, row) sequence; it is walked as segments that stay inside one
// column pair.  Both CTAs of a pair see the same segments (same rows) in adjacent strip columns 2*cp + rank.
struct Seg {
  int col, r0, r1;
};
__device__ __forceinline__ bool next_seg(int& cur, int pr_end, int h, uint32_t crank, Seg& sg) {
  if (cur >= pr_end) return false;
  const int cp = cur / h;
  sg.r0 = cur - cp * h;
  const int len = pr_end - cur < h - sg.r0 ? pr_end - cur : h - sg.r0;
  sg.r1 = sg.r0 + len;
  sg.col = 2 * cp + (int)crank;
  cur += len;
  return true;
}

template <int L>
__global__ void __launch_bounds__(THREADS, 1) dense_fused_kernel(const __grid_constant__ CUtensorMap tmap_a,
                                                                  const __grid_constant__ CUtensorMap tmap_b, const Params p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t crank = cluster_ctarank();               // 0 = leader
  const int unit = (int)(blockIdx.x >> 1);
  const int prob = p.nprob == 2 ? (unit & 1) : 0;
  const int piece = p.nprob == 2 ? (unit >> 1) : unit;
  const CUtensorMap* tmap = prob ? &tmap_b : &tmap_a;
  __nv_bfloat16* obuf = p.buf[prob];
  const int nx = p.nx, h = p.h;
  const int xrow_bytes = nx * SLAB_ROW;
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gen_base = smem_raw + (base - smem_u32(smem_raw));
  // [X ring][weights of the L layers][barriers][exchange][bias]
  const uint32_t x_base = base;
  const uint32_t w_base = base + NXR * xrow_bytes;
  const int wtotal = 3 * WT_BYTES * (L * nx + L * (L - 1));          // sum_j 3 * (nx + 2j) tiles
  const uint32_t bar_base = w_base + wtotal;
  float* xbuf = reinterpret_cast<float*>(gen_base + (bar_base - base) + BAR_BYTES);
  float* sbias = reinterpret_cast<float*>(gen_base + (bar_base - base) + BAR_BYTES + XBUF_BYTES);
  auto xfull = [&](int s) { return bar_base + 8u * s; };
  auto xempty = [&](int s) { return bar_base + 8u * (NXR + s); };
  const uint32_t w_bar = bar_base + 8u * (2 * NXR);
  const uint32_t wpeer_bar = bar_base + 8u * (2 * NXR + 1);
  auto tfull = [&](int a) { return bar_base + 8u * (2 * NXR + 2 + a); };
  auto tempty = [&](int a) { return bar_base + 8u * (2 * NXR + 2 + NACC + a); };
  // one "row stored" barrier per growth-ring slot (a single barrier per layer could be lapped: at the start of a row range a
  // layer produces several rows before its consumer's first wait, and an mbarrier only tells the last two phases apart)
  auto gready = [&](int slot_global) { return bar_base + 8u * (2 * NXR + 2 + 2 * NACC + slot_global); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * NXR + 2 + 2 * NACC + NRING_MAX);
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(gen_base + (tmem_slot - base));

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < NXR; ++s) {
      mbar_init(xfull(s), 1);
      mbar_init(xempty(s), 1);
    }
    mbar_init(w_bar, 1);
    mbar_init(wpeer_bar, 1);
    for (int a = 0; a < NACC; ++a) {
      mbar_init(tfull(a), 1);
      mbar_init(tempty(a), 16);          // 8 epilogue warps of each CTA release the leader's accumulator
    }
    for (int j = 0; j < NRING_MAX; ++j) mbar_init(gready(j), 16);
    fence_barrier_init();
    asm volatile("prefetch.tensormap [%0];" ::"l"(tmap) : "memory");
  }
  if (warp == 1) tmem_alloc2(tmem_slot, (uint32_t)TMEM_COLS);
  if (warp >= 2) {
    for (int i = (int)threadIdx.x - 64; i < L * NOUT; i += THREADS - 64) sbias[i] = __ldg(p.bias[prob][i / NOUT] + (i % NOUT));
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  pdl_launch_dependents();

  const int pr_begin = piece * p.piece_len;
  const int pr_end = pr_begin + p.piece_len < p.total_pr ? pr_begin + p.piece_len : p.total_pr;
  constexpr int DMAX = lag_of(L, L - 1);

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      mbar_expect_tx(w_bar, (uint32_t)wtotal);
      int off = 0;
      for (int j = 0; j < L; ++j) {
        const int bytes = 3 * (nx + 2 * j) * WT_BYTES;
        bulk_g2s(w_base + off, (const uint8_t*)p.wimg[prob][j] + (size_t)crank * bytes, (uint32_t)bytes, w_bar);
        off += bytes;
      }
      pdl_wait();      // weights are static; activations come from the previous kernel in the stream
      int q = 0;
      int cur = pr_begin;
      Seg sg;
      while (next_seg(cur, pr_end, h, crank, sg)) {
        const int n = sg.col < p.ncol ? sg.col / p.S : p.N;              // dummy column of an odd pair: rows past the tensor -> zeros
        const int xs = (sg.col < p.ncol ? (sg.col % p.S) : 0) * VALID - HALO;
        const int rho0 = sg.r0 - L > 0 ? sg.r0 - L : 0, rho1 = sg.r1 + L < h ? sg.r1 + L : h;
        for (int rho = rho0; rho < rho1; ++rho, ++q) {
          const int slot = q & (NXR - 1);
          mbar_wait(xempty(slot), (((uint32_t)q / NXR) & 1u) ^ 1u, p.err, 41);
          if (crank == 0) mbar_expect_tx(xfull(slot), 2u * (uint32_t)xrow_bytes);     // both CTAs' boxes land on the leader's barrier
          tma_load_4d_pair(x_base + slot * xrow_bytes, tmap, mapa_u32(xfull(slot), 0), 0, xs, n * h + rho, 0);
        }
      }
    }
  } else if (warp == 1) {
    if (crank != 0) {
      mbar_wait(w_bar, 0, p.err, 42);
      if (lane == 0) mbar_arrive_cluster(mapa_u32(wpeer_bar, 0));
    } else {
      // ===================== MMA issuer (leader): whole warp runs the loop, one elected lane issues =====================
      constexpr uint32_t idesc = umma_idesc_bf16(256, NB);
      const uint32_t hi_a = desc_hi(256, 6);             // X rows: SWIZZLE_32B, 8-row atoms of 256 bytes
      const uint32_t hi_b = desc_hi(128, 0);             // weights: no-swizzle core matrices, 8-row groups 128 bytes apart
      mbar_wait(w_bar, 0, p.err, 42);
      mbar_wait(wpeer_bar, 0, p.err, 43);
      int q_base = 0, xw = 0, gcnt = 0;
      int ev[MAXL] = {0, 0, 0, 0};        // rows of layer j whose "stored" barrier has been observed (running count)
      int prod[MAXL] = {0, 0, 0, 0};      // rows of layer j issued so far (running count): row's ring slot = count % ring
      int cur = pr_begin;
      Seg sg;
      while (next_seg(cur, pr_end, h, crank, sg)) {
        const int r0 = sg.r0, r1 = sg.r1;
        const int rho0 = r0 - L > 0 ? r0 - L : 0, rho1 = r1 + L < h ? r1 + L : h;
        int cb[MAXL];                      // running count of layer j's first row of this segment
#pragma unroll
        for (int j = 0; j < MAXL; ++j) cb[j] = prod[j];
        const int s_first = r0 - (L - 1) > 0 ? r0 - (L - 1) : 0, s_last = r1 - 1 + DMAX;
        for (int s = s_first; s <= s_last; ++s) {
          auto group = [&](auto OI) {
            constexpr int J = order_of(L, decltype(OI)::value);
            const int r = s - lag_of(L, J);
            const int lo = r0 - (L - 1 - J) > 0 ? r0 - (L - 1 - J) : 0, hi = r1 + (L - 1 - J) < h ? r1 + (L - 1 - J) : h;
            if (r < lo || r >= hi) return;
            const int acc = gcnt % NACC;
            const uint32_t use = (uint32_t)(gcnt / NACC);
            mbar_wait(tempty(acc), (use & 1u) ^ 1u, p.err, 44);
            if constexpr (J == 0) {
              const int last = r + 1 < rho1 - 1 ? r + 1 : rho1 - 1;
              const int need = q_base + (last - rho0) + 1;
              while (xw < need) {
                mbar_wait(xfull(xw & (NXR - 1)), ((uint32_t)xw / NXR) & 1u, p.err, 45);
                ++xw;
              }
            } else {
              const int lo_p = r0 - (L - J) > 0 ? r0 - (L - J) : 0, hi_p = r1 + (L - J) < h ? r1 + (L - J) : h;
              const int last = r + 1 < hi_p - 1 ? r + 1 : hi_p - 1;
              const int target = cb[J - 1] + (last - lo_p) + 1;
              while (ev[J - 1] < target) {
                constexpr int RG = ring_of(L, J - 1);
                const int c = ev[J - 1];
                mbar_wait(gready(ringslot0_of(L, J - 1) + c % RG), (uint32_t)(c / RG) & 1u, p.err, 46);
                ++ev[J - 1];
              }
            }
            tc_fence_after();
            const uint32_t d = tmem_base + (uint32_t)(acc * NB);
            const int nks = nx + 2 * J;
            const uint32_t wj = w_base + (uint32_t)(3 * WT_BYTES * (J * nx + J * (J - 1)));
            const uint32_t b_ky = (uint32_t)nks * (WT_BYTES >> 4);
            uint32_t accum = 0;
            bool ok[3];
            uint32_t xa[3];
#pragma unroll
            for (int ky = 0; ky < 3; ++ky) {
              const int rr = r + ky - 1;
              ok[ky] = rr >= 0 && rr < h;
              xa[ky] = x_base + (uint32_t)(((q_base + rr - rho0) & (NXR - 1)) * xrow_bytes);
            }
            // K order = the dense buffer's channel order [X | x1 | x2 | x3], ky inner: the order of the unfused kernel
            for (int ks = 0; ks < nx; ++ks) {
              const uint32_t b_lo = desc_lo(wj + (uint32_t)ks * WT_BYTES, (NBH / 8) * 128);
#pragma unroll
              for (int ky = 0; ky < 3; ++ky) {
                if (!ok[ky]) continue;
                const uint64_t ad = desc_join(desc_lo(xa[ky] + (uint32_t)ks * SLAB_ROW, 16), hi_a);
                const uint64_t bd = desc_join(b_lo + (uint32_t)ky * b_ky, hi_b);
                umma2_bf16_elect(d, ad, bd, idesc, accum);
                accum = 1;
              }
            }
#pragma unroll
            for (int g = 0; g < J; ++g) {
              uint32_t ga[3];
#pragma unroll
              const int lo_g = r0 - (L - 1 - g) > 0 ? r0 - (L - 1 - g) : 0;
#pragma unroll
              for (int ky = 0; ky < 3; ++ky) {
                const int cnt = ok[ky] ? cb[g] + (r + ky - 1 - lo_g) : 0;
                ga[ky] = tmem_base + (uint32_t)(ringcol_of(L, g) + (cnt % ring_of(L, g)) * 16);
              }
#pragma unroll
              for (int half = 0; half < 2; ++half) {
                const uint32_t b_lo = desc_lo(wj + (uint32_t)(nx + 2 * g + half) * WT_BYTES, (NBH / 8) * 128);
#pragma unroll
                for (int ky = 0; ky < 3; ++ky) {
                  if (!ok[ky]) continue;
                  const uint64_t bd = desc_join(b_lo + (uint32_t)ky * b_ky, hi_b);
                  umma2_ts_bf16_elect(d, ga[ky] + (uint32_t)(half * 8), bd, idesc, accum);
                  accum = 1;
                }
              }
            }
            umma2_commit_elect(tfull(acc));
            ++gcnt;
            ++prod[J];
          };
          group(IC<0>{});
          group(IC<1>{});
          group(IC<2>{});
          if constexpr (L > 3) group(IC<3>{});
          // X row s - DMAX - 1 has no reader left
          const int f = s - DMAX - 1;
          if (f >= rho0 && f < rho1) umma2_commit_elect(xempty((q_base + f - rho0) & (NXR - 1)));
        }
        for (int f = (r1 - 1 > rho0 ? r1 - 1 : rho0); f < rho1; ++f) umma2_commit_elect(xempty((q_base + f - rho0) & (NXR - 1)));
        q_base += rho1 - rho0;
      }
    }
  } else {
    // ===================== epilogue warps 2..9: two warpgroups x four lane quarters =====================
    const int wg = (warp - 2) >> 2;              // output channels 16 wg .. 16 wg + 15
    const int q = warp & 3;                      // TMEM lane quarter = positions 32 q .. 32 q + 31 of the strip row
    const int i = q * 32 + lane;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16);
    const uint32_t tempty_leader0 = mapa_u32(tempty(0), 0);
    const uint32_t gready_leader0 = mapa_u32(gready(0), 0);
    const size_t slab_elems = (size_t)p.slabM * 16;
    pdl_wait();
    int gcnt = 0;
    int cnt[MAXL] = {0, 0, 0, 0};         // rows of layer j stored so far: ring slot = count % ring (the MMA warp counts the same)
    int cur = pr_begin;
    Seg sg;
    while (next_seg(cur, pr_end, h, crank, sg)) {
      const int r0 = sg.r0, r1 = sg.r1;
      const bool col_ok = sg.col < p.ncol;
      const int n = col_ok ? sg.col / p.S : 0;
      const int x = (col_ok ? (sg.col % p.S) : 0) * VALID - HALO + i;
      const bool inimg = col_ok && x >= 0 && x < p.w;
      const bool store_col = inimg && i >= HALO && i < MPOS - HALO;
      const int s_first = r0 - (L - 1) > 0 ? r0 - (L - 1) : 0, s_last = r1 - 1 + DMAX;
      for (int s = s_first; s <= s_last; ++s) {
        auto group = [&](auto OI) {
          constexpr int J = order_of(L, decltype(OI)::value);
          const int r = s - lag_of(L, J);
          const int lo = r0 - (L - 1 - J) > 0 ? r0 - (L - 1 - J) : 0, hi = r1 + (L - 1 - J) < h ? r1 + (L - 1 - J) : h;
          if (r < lo || r >= hi) return;
          const int acc = gcnt % NACC;
          const uint32_t use = (uint32_t)(gcnt / NACC);
          mbar_wait(tfull(acc), use & 1u, p.err, 47);
          tc_fence_after();
          uint32_t a0[16], a1[16], a2[16];
          const uint32_t trow = lane_addr + (uint32_t)(acc * NB + wg * 16);
          tmem_ld16(trow, a0);
          tmem_ld16(trow + NOUT, a1);
          tmem_ld16(trow + 2 * NOUT, a2);
          tmem_ld_wait();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_cluster(tempty_leader0 + 8u * (uint32_t)acc);
          // neighbours across the quarter boundaries: lane 31's kx=0 partials go right, lane 0's kx=2 partials go left
          float* xb = xbuf + (size_t)((wg * 2 + (gcnt & 1)) * 4) * 32;
          if (lane == 31) {
#pragma unroll
            for (int t = 0; t < 16; ++t) xb[(q * 2 + 0) * 16 + t] = __uint_as_float(a0[t]);
          }
          if (lane == 0) {
#pragma unroll
            for (int t = 0; t < 16; ++t) xb[(q * 2 + 1) * 16 + t] = __uint_as_float(a2[t]);
          }
          named_bar_sync(1 + wg, 128);
          float v[16];
#pragma unroll
          for (int t = 0; t < 16; ++t) {
            float left = __shfl_up_sync(0xffffffffu, __uint_as_float(a0[t]), 1);
            float right = __shfl_down_sync(0xffffffffu, __uint_as_float(a2[t]), 1);
            if (lane == 0) left = q > 0 ? xb[((q - 1) * 2 + 0) * 16 + t] : 0.f;
            if (lane == 31) right = q < 3 ? xb[((q + 1) * 2 + 1) * 16 + t] : 0.f;
            const float o = lrelu02(left + __uint_as_float(a1[t]) + right + sbias[J * NOUT + wg * 16 + t]);
            v[t] = inimg ? o : 0.f;            // columns outside the image are the next layer's zero padding
          }
          uint32_t pk[8];
#pragma unroll
          for (int t = 0; t < 8; ++t) pk[t] = pack_bf2(v[2 * t], v[2 * t + 1]);
          if constexpr (J < L - 1) {
            const int slot = cnt[J] % ring_of(L, J);
            tmem_st8(lane_addr + (uint32_t)(ringcol_of(L, J) + slot * 16 + wg * 8), pk);
            tmem_st_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(gready_leader0 + 8u * (uint32_t)(ringslot0_of(L, J) + slot));
            ++cnt[J];
          }
          if (store_col && r >= r0 && r < r1) {
            __nv_bfloat16* o = obuf + (size_t)(nx + 2 * J + wg) * slab_elems + ((size_t)((size_t)n * h + r) * p.w + x) * 16;
            *reinterpret_cast<uint4*>(o) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
            *reinterpret_cast<uint4*>(o + 8) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
          }
          ++gcnt;
        };
        group(IC<0>{});
        group(IC<1>{});
        group(IC<2>{});
        if constexpr (L > 3) group(IC<3>{});
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();      // neither CTA leaves (or frees tensor memory) while the other may still signal it
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc2(tmem_base, (uint32_t)TMEM_COLS);
  }
}

static int smem_bytes(int L, int nx) {
  return 1024 + NXR * nx * SLAB_ROW + 3 * WT_BYTES * (L * nx + L * (L - 1)) + BAR_BYTES + XBUF_BYTES + BIAS_BYTES;
}

}  // namespace dbf

bool dense_fused_supported(int cin, int L) {
  if (cin % 16 != 0 || (L != 3 && L != 4)) return false;
  return dbf::smem_bytes(L, cin / 16) <= 227 * 1024;
}

int launch_dense_fused(const TcConvW* w, int L, __nv_bfloat16* buf, long long slabM, int cin, int N, int h, int wd, cudaStream_t st,
                       const TcConvW* w2, __nv_bfloat16* buf2) {
  SELFC_CHECK_ARG(dense_fused_supported(cin, L), "dense_fused: cin=%d with %d fused layers does not fit shared memory", cin, L);
  SELFC_CHECK_ARG(aligned16(buf) && slabM == (long long)N * h * wd, "dense_fused: slab layout / alignment");
  const bool dual = w2 != nullptr;
  SELFC_CHECK_ARG(!dual || (buf2 != nullptr && aligned16(buf2) && buf2 != buf), "dense_fused: the second problem needs its own buffer");
  const int nx = cin / 16;
  for (int j = 0; j < L; ++j) {
    SELFC_CHECK_ARG(w[j].img_pair != nullptr && w[j].cin_buf == cin + 32 * j, "dense_fused: layer %d weights not packed for cin=%d", j, cin + 32 * j);
    SELFC_CHECK_ARG(!dual || (w2[j].img_pair != nullptr && w2[j].cin_buf == cin + 32 * j), "dense_fused: second problem's layer %d", j);
  }
  tc::EncodeTiledFn encode = tc::get_encode_fn();
  if (!encode) {
    set_error("cuTensorMapEncodeTiled is not available from this driver");
    return SELFC_E_CUDA;
  }
  // X rows: {16 channels, W positions, N*h rows (frames stacked: out-of-image rows are skipped, never fetched), nx slabs}
  const cuuint64_t gdim[4] = {16, (cuuint64_t)wd, (cuuint64_t)N * h, (cuuint64_t)nx};
  const cuuint64_t gstr[3] = {32, (cuuint64_t)wd * 32, (cuuint64_t)slabM * 32};
  const cuuint32_t box[4] = {16, (cuuint32_t)dbf::MPOS, 1, (cuuint32_t)nx};
  const cuuint32_t estr[4] = {1, 1, 1, 1};
  CUtensorMap tmap, tmap2;
  CUresult r = encode(&tmap, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, buf, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                      CU_TENSOR_MAP_SWIZZLE_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r == CUDA_SUCCESS)
    r = encode(&tmap2, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, dual ? buf2 : buf, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("dense_fused: cuTensorMapEncodeTiled failed with CUresult %d (%dx%dx%d, %d slabs)", (int)r, N, h, wd, nx);
    return SELFC_E_CUDA;
  }
  dbf::Params p;
  memset(&p, 0, sizeof(p));
  for (int j = 0; j < L; ++j) {
    p.wimg[0][j] = w[j].img_pair;
    p.bias[0][j] = w[j].bias;
    p.wimg[1][j] = dual ? w2[j].img_pair : w[j].img_pair;
    p.bias[1][j] = dual ? w2[j].bias : w[j].bias;
  }
  p.buf[0] = buf;
  p.buf[1] = dual ? buf2 : buf;
  p.nprob = dual ? 2 : 1;
  p.nx = nx;
  p.slabM = slabM;
  p.N = N;
  p.h = h;
  p.w = wd;
  p.S = cdiv(wd, dbf::VALID);
  p.ncol = N * p.S;
  p.total_pr = cdiv(p.ncol, 2) * h;
  if (p.total_pr == 0) return 0;
  const int pairs_avail = tc::num_sms() / 2 / p.nprob;
  SELFC_CHECK_ARG(pairs_avail >= 1, "dense_fused: needs at least %d SMs", 2 * p.nprob);
  // every piece pays 2 (L - 1) rows of recompute where it starts / ends: do not cut tiny clips into one-row pieces
  int piece_len = cdiv(p.total_pr, pairs_avail);
  if (piece_len < 8) piece_len = 8;
  p.piece_len = piece_len;
  const int pieces = cdiv(p.total_pr, piece_len);
  p.err = tc::err_flag_for_device();
  const int smem = dbf::smem_bytes(L, nx);
  static bool smem_set[64] = {};           // per device: function attributes belong to the device's context
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev >= 0 && dev < 64 && !smem_set[dev]) {
    SELFC_CUDA(cudaFuncSetAttribute(dbf::dense_fused_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    SELFC_CUDA(cudaFuncSetAttribute(dbf::dense_fused_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    smem_set[dev] = true;
  }
  const int grid = 2 * pieces * p.nprob;
  if (L == 4) SELFC_CUDA(tc::launch_pdl_pairs(dbf::dense_fused_kernel<4>, grid, dbf::THREADS, smem, st, tmap, tmap2, p));
  else SELFC_CUDA(tc::launch_pdl_pairs(dbf::dense_fused_kernel<3>, grid, dbf::THREADS, smem, st, tmap, tmap2, p));
  SELFC_LAUNCH_CHECK("dense_fused_kernel");
  return 0;
}

}  // namespace selfc
