// conv1..convL of a D2DTInput dense block (Subnet_constructor.py:102-105,126-129) as ONE launch: the growth channels
// x1..x_{L-1} never leave the SM between the layers.  BF16 mode, tcgen05 / TMEM / TMA, CTA pairs (cta_group::2).
//
// Layer by layer (conv_tc3.cu) every layer re-reads [X | x1 | ...] from HBM: (cin + 32(k-1)) channels in, 32 out, per layer.  Here
// a CTA owns a vertical STRIP of the image -- 128 positions wide, 120 of them outputs of the last layer -- and slides down it one
// row per step with the layers software-pipelined one behind the other ("line buffers"); e.g. for the one-slab-X blocks
//
//     step s issues   conv1 row s,   conv2 row s-2,   conv3 row s-4,   conv4 row s-6
//
// (the schedules -- lags, issue order, ring sizes -- are the tables below; C-ABI selfc_dense_fused_schedule exposes them to a CPU test
// that simulates the issue order).  A row of a layer is one accumulator: D[p][kx*32+n] += sum_c A[row+ky-1][p][c] * W[ky,kx][n][c]
// (M = 128 positions of this CTA + 128 of its pair, N = 96 = the three kx taps stacked, K = 16), out[p][n] = D[p-1][kx=0] +
// D[p][kx=1] + D[p+1][kx=2] -- the same MMAs in the same K order and the same epilogue arithmetic as conv3x3_tc3_kernel, so the
// results are bit-identical to the layer-by-layer path.
// The A operands: X rows come from HBM once, by TMA (SWIZZLE_32B, one box per row: 128 positions x nx slabs), into a ring in shared
// memory; the growth rows are written by the epilogue warps as bf16 straight into TENSOR MEMORY (tcgen05.st) and read from there by
// the MMAs (A-in-TMEM form: lane = position, 8 columns per 16 channels; 48 cycles per N = 96 MMA, profiles/r2b_ubench_mma_tmem_a.txt)
// -- rings of 8 + 6 + 4 rows next to two 96-column accumulators.  No vertical recompute (only 2 (L-1) rows where a CTA's row range
// starts / ends), horizontal recompute 8 of 128 positions (the 15/16 efficiency of the 30-of-32 tiles of the layer-by-layer kernel).
// Out-of-image rows are MMAs not issued, out-of-image columns are TMA zero-fill (X) or zeros written by the epilogue (growth rows):
// exactly the zero padding of the reference.  Every layer's 32 outputs are stored to the dense buffer in HBM once (conv5 reads
// them): per pixel a block moves (cin + 128) bf16 channels instead of (4 cin + 192 + 128) -- and for the F block (3 outputs) not
// even that: conv5's three temporal taps are applied to the row while it is on chip (schedule 3) and 9 fp32 partial products per
// pixel are all that leaves; f5_combine_kernel finishes conv5 and the additive coupling.
//
// Per CTA (576 threads): warp 0 TMA producer, warp 1 MMA issuer (leader CTA only; interior rows take an unpredicated issue path
// that stays in the uniform datapath), warps 2..17 epilogue = two teams x two channel halves x four lane quarters: team t drains the
// groups with odd / even running index = accumulator t, so two rows' epilogues are in flight.  The lane +-1 neighbours of the kx
// combination are warp shuffles + a 512-byte exchange across the lane quarters, or -- where shared memory allows (one-slab X) -- all
// partial sums through shared memory.  A "row stored" mbarrier per ring slot tells the issuer when a growth row may be read.
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
constexpr int NACC = 2;                       // accumulators: group g uses g & 1, and is drained by epilogue team g & 1
constexpr int MAXL = 4;
constexpr int EPI_WARPS = 16;                 // two teams x two channel halves x four lane quarters
constexpr int THREADS = 64 + 32 * EPI_WARPS;
constexpr int TMEM_COLS = 512;
constexpr int RING_COL0 = NACC * NB;          // 192
constexpr int BAR_BYTES = 512;
constexpr int NXR_MAX = 16;
constexpr int NRING_MAX = 18;                 // growth-ring slots of all layers (8 + 6 + 4)
constexpr int XBUF_BYTES = 2 * 2 * 2 * 4 * 2 * 16 * 4;      // [team][channel half][parity][quarter][side][16] floats
constexpr int BIAS_BYTES = MAXL * NOUT * 4;
// SCH 0 (X of one slab) has shared memory to spare: there the kx partial sums of ALL lanes travel through shared memory (8 STS.128
// + 8 LDS.128 per thread and row) instead of 32 shuffles + 32 edge selects -- the epilogue is instruction-issue bound.  Rows of
// 20 floats (16 + 4 pad: conflict-free 16-byte accesses), row 0 of L / row 128 of R stay zero (the strip's ends have no neighbour).
constexpr int XS_ROW = 20;
constexpr int XS_BYTES_PER = 2 * (MPOS + 1) * XS_ROW * 4;       // L and R planes of one (team, channel half)
__host__ __device__ constexpr bool xs_of(int SCH) { return SCH == 0; }

// Row schedules (SCH): which row of which layer a step issues, in which order, and the ring sizes that follow from it (a ring slot
// may be overwritten once every reader of its row has been ISSUED before the overwriting row's own MMAs; verified by simulation).
//   SCH 0  L = 4, lags 0,2,4,6, order conv1,conv2,conv3,conv4: every consumer reads rows stored a whole step earlier.  Rings
//          8 + 6 + 4 rows, 9 live X rows (16-slot ring): X of 1 slab (G, H, local_m1).
//   SCH 1  L = 4, lags 0,2,4,5, order conv3,conv1,conv2,conv4: conv4 reads the x3 row conv3 stored two groups earlier in the same
//          step.  Rings 7 + 5 + 3, 8 live X rows (9-slot ring): what fits next to 108 KB of weights when X has 3 slabs (F).
//   SCH 2  L = 3, lags 0,2,4: rings 6 + 4, 7 live X rows (8-slot ring): X of 4 slabs (STP 64 -> 64; conv4 runs layer-by-layer).
//   SCH 3  SCH 1 plus a FIFTH group per step: the three temporal taps of conv5 applied to row s-6 of [X | x1..x4] (a pointwise
//          GEMM, N = 16 = 3 taps x 3 outputs padded), order conv3, conv1, conv5-taps, conv2, conv4; x4 gets a 2-row ring.  Only
//          for cout = 3 (the F block of a coupling): 9 fp32 partial products per pixel replace conv5's re-read of 176 channels,
//          and x1..x4 need not be written to HBM at all.
__host__ __device__ constexpr bool f5_of(int SCH) { return SCH == 3; }
__host__ __device__ constexpr int nlayers_of(int SCH) { return SCH == 2 ? 3 : 4; }
__host__ __device__ constexpr int lag_of(int SCH, int j) { return (SCH == 1 || SCH == 3) ? (j == 4 ? 6 : j == 3 ? 5 : 2 * j) : 2 * j; }
__host__ __device__ constexpr int order_of(int SCH, int oi) {
  return SCH == 1 ? (oi == 0 ? 2 : oi == 1 ? 0 : oi == 2 ? 1 : 3) : SCH == 3 ? (oi == 0 ? 2 : oi == 1 ? 0 : oi == 2 ? 4 : oi == 3 ? 1 : 3) : oi;
}
__host__ __device__ constexpr int ngroups_of(int SCH) { return nlayers_of(SCH) + (f5_of(SCH) ? 1 : 0); }
__host__ __device__ constexpr int ring_of(int SCH, int j) { return SCH == 0 ? 8 - 2 * j : (SCH == 1 || SCH == 3) ? (j == 3 ? 2 : 7 - 2 * j) : 6 - 2 * j; }
__host__ __device__ constexpr int nxr_of(int SCH) { return SCH == 0 ? 16 : (SCH == 1 || SCH == 3) ? 9 : 8; }
__host__ __device__ constexpr int ringslot0_of(int SCH, int j) {
  return j == 0 ? 0 : j == 1 ? ring_of(SCH, 0) : j == 2 ? ring_of(SCH, 0) + ring_of(SCH, 1) : ring_of(SCH, 0) + ring_of(SCH, 1) + ring_of(SCH, 2);
}
constexpr int F5_N = 16;                                  // conv5-taps GEMM: rows tap * 3 + co (9 of 16)
constexpr int F5_KSTEPS = 11;                             // 48 + 128 channels
constexpr int F5_TILE = (F5_N / 2) * 16 * 2;              // one K-step of one CTA's half of the B image: 256 bytes
__host__ __device__ constexpr int ringcol_of(int SCH, int j) { return RING_COL0 + 16 * ringslot0_of(SCH, j); }
static_assert(ringcol_of(0, 2) + 16 * ring_of(0, 2) <= TMEM_COLS, "TMEM budget");
static_assert(ringslot0_of(0, 2) + ring_of(0, 2) <= NRING_MAX, "ring slot barriers");
static_assert(ringcol_of(3, 3) + 16 * ring_of(3, 3) <= TMEM_COLS && ringslot0_of(3, 3) + ring_of(3, 3) <= NRING_MAX, "SCH 3 budget");

struct Params {
  const void* wimg[2][MAXL];     // per problem, per layer: TcConvW::img_pair (two halves of the B image)
  const float* bias[2][MAXL];
  __nv_bfloat16* buf[2];
  int nprob, nx;                 // nx = 16-channel slabs of X
  long long slabM;
  int N, h, w, S, ncol;          // S strips per image row, ncol = N * S strip columns
  int piece_len, total_pr;       // a CTA pair owns piece_len consecutive rows of the (column pair, row) sequence
  const void* w5img;             // SCH 3: conv5-taps B image (pack_f5_kernel), two halves
  float* part;                   // SCH 3: partial products [3 taps][M][4] fp32 (M = N*h*w)
  int store_growth;              // write x1..xL to the dense buffer (0: nobody reads them -- SCH 3)
  int* err;
  long long* dbg;                // SELFC_TC_DBG=1: CTA 0's barrier-wait cycles per role (tc::debug_next_slot)
};

__device__ __forceinline__ void tma_load_4d_pair(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
// Pair MMAs issued by one elected lane when `valid` (a warp-uniform flag): skipping an out-of-image tap must not put a branch
// around the instruction -- with one basic block per MMA the compiler re-materialises every descriptor through R2UR moves
// (~90 cycles per issue instead of a handful of uniform-datapath instructions).
__device__ __forceinline__ void umma2_ss_bf16_elect_if(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate,
                                                       uint32_t valid) {
  asm volatile(
      "{\n\t.reg .pred p, q, v;\n\t"
      "elect.sync _|q, 0xffffffff;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "setp.ne.b32 v, %5, 0;\n\t"
      "and.pred q, q, v;\n\t"
      "@q tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(valid)
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
__device__ __forceinline__ void umma2_ts_bf16_elect_if(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate,
                                                       uint32_t valid) {
  asm volatile(
      "{\n\t.reg .pred p, q, v;\n\t"
      "elect.sync _|q, 0xffffffff;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "setp.ne.b32 v, %5, 0;\n\t"
      "and.pred q, q, v;\n\t"
      "@q tcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(valid)
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

// Barrier waits of this kernel are on the critical path of every row (producer -> epilogue -> consumer), unlike the tile kernels
// whose waits mostly find the phase complete: poll without a suspend-time hint (the hinted form parks the warp and wakes late).
__device__ __forceinline__ bool mbar_try_wait_nohint(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait_fast(uint32_t bar, uint32_t parity, int* err, int code) {
#ifdef SELFC_DBF_HINT_WAIT
  mbar_wait(bar, parity, err, code);
  return;
#endif
  uint32_t spins = 0;
  while (!mbar_try_wait_nohint(bar, parity)) {
    if (++spins > (1u << 26)) {       // a protocol bug traps (launch error) instead of hanging the GPU
      if (err) atomicExch(err, code);
      __trap();
    }
  }
}

// the same, accumulating the cycles spent waiting when the debug counters are on
__device__ __forceinline__ void mbar_wait_t(uint32_t bar, uint32_t parity, int* err, int code, bool timed, long long& acc) {
  if (timed) {
    const long long t0 = clock64();
    mbar_wait_fast(bar, parity, err, code);
    acc += clock64() - t0;
  } else {
    mbar_wait_fast(bar, parity, err, code);
  }
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

template <int SCH, int NX, bool DBG>
__global__ void __launch_bounds__(THREADS, 1) dense_fused_kernel(const __grid_constant__ CUtensorMap tmap_a,
                                                                  const __grid_constant__ CUtensorMap tmap_b, const Params p) {
  constexpr int L = nlayers_of(SCH);
  constexpr int NXR = nxr_of(SCH);
  constexpr bool F5 = f5_of(SCH);
  constexpr int DMAX = lag_of(SCH, F5 ? 4 : L - 1);
  // oldest X row a step reads = s - XOLD: a conv layer at lag d reads rows down to s - d - 1, the conv5-taps group only row s - 6
  constexpr int XOLD = F5 ? (lag_of(SCH, L - 1) + 1 > lag_of(SCH, 4) ? lag_of(SCH, L - 1) + 1 : lag_of(SCH, 4)) : lag_of(SCH, L - 1) + 1;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t crank = cluster_ctarank();               // 0 = leader
  const int unit = (int)(blockIdx.x >> 1);
  const int prob = p.nprob == 2 ? (unit & 1) : 0;
  const int piece = p.nprob == 2 ? (unit >> 1) : unit;
  const CUtensorMap* tmap = prob ? &tmap_b : &tmap_a;
  __nv_bfloat16* obuf = p.buf[prob];
  constexpr int nx = NX;                                  // 16-channel slabs of X: compile-time, so every descriptor offset folds
  const int h = p.h;
  constexpr int xrow_bytes = nx * SLAB_ROW;
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gen_base = smem_raw + (base - smem_u32(smem_raw));
  // [X ring][weights of the L layers][barriers][exchange][bias]
  const uint32_t x_base = base;
  const uint32_t w_base = base + NXR * xrow_bytes;
  constexpr int wconv = 3 * WT_BYTES * (L * nx + L * (L - 1));       // sum_j 3 * (nx + 2j) tiles
  constexpr int wtotal = wconv + (F5 ? F5_KSTEPS * F5_TILE : 0);
  const uint32_t bar_base = w_base + wtotal;
  float* xbuf = reinterpret_cast<float*>(gen_base + (bar_base - base) + BAR_BYTES);
  float* sbias = reinterpret_cast<float*>(gen_base + (bar_base - base) + BAR_BYTES + XBUF_BYTES);
  float* xsbuf = reinterpret_cast<float*>(gen_base + (bar_base - base) + BAR_BYTES + XBUF_BYTES + BIAS_BYTES);     // xs_of(SCH) only
  auto xfull = [&](int s) { return bar_base + 8u * s; };
  auto xempty = [&](int s) { return bar_base + 8u * (NXR_MAX + s); };
  const uint32_t w_bar = bar_base + 8u * (2 * NXR_MAX);
  const uint32_t wpeer_bar = bar_base + 8u * (2 * NXR_MAX + 1);
  auto tfull = [&](int a) { return bar_base + 8u * (2 * NXR_MAX + 2 + a); };
  auto tempty = [&](int a) { return bar_base + 8u * (2 * NXR_MAX + 2 + NACC + a); };
  // one "row stored" barrier per growth-ring slot (a single barrier per layer could be lapped: at the start of a row range a
  // layer produces several rows before its consumer's first wait, and an mbarrier only tells the last two phases apart)
  auto gready = [&](int slot_global) { return bar_base + 8u * (2 * NXR_MAX + 2 + 2 * NACC + slot_global); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * NXR_MAX + 2 + 2 * NACC + NRING_MAX);
  static_assert(8 * (2 * NXR_MAX + 2 + 2 * NACC + NRING_MAX + 1) <= BAR_BYTES, "barrier block");
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
      mbar_init(tempty(a), 16);          // the 8 warps of one epilogue team in each CTA release the leader's accumulator
    }
    for (int j = 0; j < NRING_MAX; ++j) mbar_init(gready(j), 16);
    fence_barrier_init();
    asm volatile("prefetch.tensormap [%0];" ::"l"(tmap) : "memory");
  }
  if (warp == 1) tmem_alloc2(tmem_slot, (uint32_t)TMEM_COLS);
  if (warp >= 2) {
    for (int i = (int)threadIdx.x - 64; i < L * NOUT; i += THREADS - 64) sbias[i] = __ldg(p.bias[prob][i / NOUT] + (i % NOUT));
    if constexpr (xs_of(SCH)) {
      for (int i = (int)threadIdx.x - 64; i < 4 * XS_BYTES_PER / 4; i += THREADS - 64) xsbuf[i] = 0.f;
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  pdl_launch_dependents();

  const int pr_begin = piece * p.piece_len;
  const int pr_end = pr_begin + p.piece_len < p.total_pr ? pr_begin + p.piece_len : p.total_pr;
  const bool timed = DBG && p.dbg != nullptr && blockIdx.x == 0;

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
      if constexpr (F5)
        bulk_g2s(w_base + wconv, (const uint8_t*)p.w5img + (size_t)crank * (F5_KSTEPS * F5_TILE), (uint32_t)(F5_KSTEPS * F5_TILE), w_bar);
      pdl_wait();      // weights are static; activations come from the previous kernel in the stream
      int slot = 0;
      uint32_t ph = 0;
      int cur = pr_begin;
      Seg sg;
      long long w_xempty = 0;
      const long long t_start = DBG ? clock64() : 0;
      while (next_seg(cur, pr_end, h, crank, sg)) {
        const int n = sg.col < p.ncol ? sg.col / p.S : p.N;              // dummy column of an odd pair: rows past the tensor -> zeros
        const int xs = (sg.col < p.ncol ? (sg.col % p.S) : 0) * VALID - HALO;
        const int rho0 = sg.r0 - L > 0 ? sg.r0 - L : 0, rho1 = sg.r1 + L < h ? sg.r1 + L : h;
        for (int rho = rho0; rho < rho1; ++rho) {
          mbar_wait_t(xempty(slot), ph ^ 1u, p.err, 41, timed, w_xempty);
          if (crank == 0) mbar_expect_tx(xfull(slot), 2u * (uint32_t)xrow_bytes);     // both CTAs' boxes land on the leader's barrier
          tma_load_4d_pair(x_base + slot * xrow_bytes, tmap, mapa_u32(xfull(slot), 0), 0, xs, n * h + rho, 0);
          if (++slot == NXR) { slot = 0; ph ^= 1u; }
        }
      }
      if (timed) { p.dbg[0] = w_xempty; p.dbg[1] = clock64() - t_start; }
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
      const uint32_t xlo_base = desc_lo(x_base, 16);                    // descriptor low words: + (byte offset >> 4)
      const uint32_t wlo_base = desc_lo(w_base, (NBH / 8) * 128);
      int q_base = 0, xw = 0, gcnt = 0;   // q_base / xw: X rows loaded before this segment / observed (running counts)
      int ev[MAXL] = {0, 0, 0, 0};        // rows of layer j whose "stored" barrier has been observed (running count)
      int prod[MAXL] = {0, 0, 0, 0};      // rows of layer j issued so far (running count): row's ring slot = count % ring
      int cur = pr_begin;
      Seg sg;
      long long w_tempty = 0, w_xfull = 0, w_gready[MAXL] = {0, 0, 0, 0};
      int nsteps = 0;
      const long long t_start = DBG ? clock64() : 0;
      while (next_seg(cur, pr_end, h, crank, sg)) {
        const int r0 = sg.r0, r1 = sg.r1;
        const int rho0 = r0 - L > 0 ? r0 - L : 0, rho1 = r1 + L < h ? r1 + L : h;
        int cb[MAXL];                      // running count of layer j's first row of this segment
#pragma unroll
        for (int j = 0; j < MAXL; ++j) cb[j] = prod[j];
        const int s_first = r0 - (L - 1) > 0 ? r0 - (L - 1) : 0, s_last = r1 - 1 + DMAX;
        nsteps += s_last - s_first + 1;
        for (int s = s_first; s <= s_last; ++s) {
          auto group = [&](auto JC) {
            constexpr int J = decltype(JC)::value;
            const int r = s - lag_of(SCH, J);
            const int lo = r0 - (L - 1 - J) > 0 ? r0 - (L - 1 - J) : 0, hi = r1 + (L - 1 - J) < h ? r1 + (L - 1 - J) : h;
            if (r < lo || r >= hi) return;
            const int acc = gcnt & 1;
            const uint32_t use = (uint32_t)(gcnt >> 1);
            mbar_wait_t(tempty(acc), (use & 1u) ^ 1u, p.err, 44, timed, w_tempty);
            if constexpr (J == 0) {
              const int last = r + 1 < rho1 - 1 ? r + 1 : rho1 - 1;
              const int need = q_base + (last - rho0) + 1;
              while (xw < need) {
                mbar_wait_t(xfull(xw % NXR), ((uint32_t)xw / NXR) & 1u, p.err, 45, timed, w_xfull);
                ++xw;
              }
            } else {
              const int lo_p = r0 - (L - J) > 0 ? r0 - (L - J) : 0, hi_p = r1 + (L - J) < h ? r1 + (L - J) : h;
              const int last = r + 1 < hi_p - 1 ? r + 1 : hi_p - 1;
              const int target = cb[J - 1] + (last - lo_p) + 1;
              while (ev[J - 1] < target) {
                constexpr int RG = ring_of(SCH, J - 1);
                const int c = ev[J - 1];
                mbar_wait_t(gready(ringslot0_of(SCH, J - 1) + c % RG), (uint32_t)(c / RG) & 1u, p.err, 46, timed, w_gready[J - 1]);
                ++ev[J - 1];
              }
            }
            tc_fence_after();
            const uint32_t d = tmem_base + (uint32_t)(acc * NB);
            constexpr int nks = nx + 2 * J;
            constexpr uint32_t wj_off = (uint32_t)(3 * WT_BYTES * (J * nx + J * (J - 1)));      // this layer's B image inside the weights
            constexpr uint32_t b_ky = (uint32_t)nks * (WT_BYTES >> 4);
            // slots of rows r-1, r, r+1 in the X ring / the growth rings: one modulo for the middle row, wrap for its neighbours
            const bool ok_up = r > 0, ok_dn = r + 1 < h;                  // out-of-image rows are skipped taps (zero padding)
            uint32_t xa[3];
            {
              const int sm = (q_base + r - rho0) % NXR;
              const int su = sm == 0 ? NXR - 1 : sm - 1, sd = sm == NXR - 1 ? 0 : sm + 1;
              xa[0] = xlo_base + (uint32_t)(su * (xrow_bytes >> 4));
              xa[1] = xlo_base + (uint32_t)(sm * (xrow_bytes >> 4));
              xa[2] = xlo_base + (uint32_t)(sd * (xrow_bytes >> 4));
            }
            uint32_t ga[J > 0 ? J : 1][3];
#pragma unroll
            for (int g = 0; g < J; ++g) {
              const int RGg = ring_of(SCH, g);
              const int lo_g = r0 - (L - 1 - g) > 0 ? r0 - (L - 1 - g) : 0;
              const int sm = (cb[g] + (r - lo_g)) % RGg;
              const int su = sm == 0 ? RGg - 1 : sm - 1, sd = sm == RGg - 1 ? 0 : sm + 1;
              const uint32_t gb = tmem_base + (uint32_t)ringcol_of(SCH, g);
              ga[g][0] = gb + (uint32_t)(su * 16);
              ga[g][1] = gb + (uint32_t)(sm * 16);
              ga[g][2] = gb + (uint32_t)(sd * 16);
            }
            // K order = the dense buffer's channel order [X | x1 | x2 | x3], ky inner: the order of the unfused kernel
            if (ok_up && ok_dn) {
              // interior row (all but the first / last image row): unpredicated issue, a few uniform-datapath instructions per MMA
#pragma unroll
              for (int ks = 0; ks < nx; ++ks) {
#pragma unroll
                for (int ky = 0; ky < 3; ++ky) {
                  const uint64_t ad = desc_join(xa[ky] + (uint32_t)(ks * (SLAB_ROW >> 4)), hi_a);
                  const uint64_t bd = desc_join(wlo_base + ((wj_off + (uint32_t)ks * WT_BYTES) >> 4) + (uint32_t)ky * b_ky, hi_b);
                  umma2_bf16_elect(d, ad, bd, idesc, (ks > 0 || ky > 0) ? 1u : 0u);
                }
              }
#pragma unroll
              for (int g = 0; g < J; ++g) {
#pragma unroll
                for (int half = 0; half < 2; ++half) {
#pragma unroll
                  for (int ky = 0; ky < 3; ++ky) {
                    const uint64_t bd = desc_join(wlo_base + ((wj_off + (uint32_t)(nx + 2 * g + half) * WT_BYTES) >> 4) + (uint32_t)ky * b_ky, hi_b);
                    umma2_ts_bf16_elect(d, ga[g][ky] + (uint32_t)(half * 8), bd, idesc, 1u);
                  }
                }
              }
            } else {
              // first / last image row: the missing tap is predicated off (never a branch around an MMA: see umma2_*_elect_if);
              // the accumulator is overwritten by the first tap issued -- (K-step 0, ky = 0), or ky = 1 when row r - 1 does not exist
              const uint32_t ok[3] = {ok_up ? 1u : 0u, 1u, ok_dn ? 1u : 0u};
              const uint32_t acc_k0[3] = {0u, ok[0], 1u};
#pragma unroll
              for (int ks = 0; ks < nx; ++ks) {
#pragma unroll
                for (int ky = 0; ky < 3; ++ky) {
                  const uint64_t ad = desc_join(xa[ky] + (uint32_t)(ks * (SLAB_ROW >> 4)), hi_a);
                  const uint64_t bd = desc_join(wlo_base + ((wj_off + (uint32_t)ks * WT_BYTES) >> 4) + (uint32_t)ky * b_ky, hi_b);
                  umma2_ss_bf16_elect_if(d, ad, bd, idesc, ks == 0 ? acc_k0[ky] : 1u, ok[ky]);
                }
              }
#pragma unroll
              for (int g = 0; g < J; ++g) {
#pragma unroll
                for (int half = 0; half < 2; ++half) {
#pragma unroll
                  for (int ky = 0; ky < 3; ++ky) {
                    const uint64_t bd = desc_join(wlo_base + ((wj_off + (uint32_t)(nx + 2 * g + half) * WT_BYTES) >> 4) + (uint32_t)ky * b_ky, hi_b);
                    umma2_ts_bf16_elect_if(d, ga[g][ky] + (uint32_t)(half * 8), bd, idesc, 1u, ok[ky]);
                  }
                }
              }
            }
            umma2_commit_elect(tfull(acc));
            ++gcnt;
            ++prod[J];
          };
          // SCH 3: the three temporal taps of conv5 on row r of [X | x1 | x2 | x3 | x4] -- a pointwise GEMM, N = 16, 11 K-steps
          auto group5 = [&]() {
            const int r = s - lag_of(SCH, 4);
            if (r < r0 || r >= r1) return;
            const int acc = gcnt & 1;
            const uint32_t use = (uint32_t)(gcnt >> 1);
            mbar_wait_t(tempty(acc), (use & 1u) ^ 1u, p.err, 44, timed, w_tempty);
            {
              // x4 row r stored?  (X and x1..x3 of row r were waited for by conv4's row r - 1 at the latest)
              constexpr int RG = ring_of(SCH, 3);
              const int target = cb[3] + (r - r0) + 1;
              while (ev[3] < target) {
                const int c = ev[3];
                mbar_wait_t(gready(ringslot0_of(SCH, 3) + c % RG), (uint32_t)(c / RG) & 1u, p.err, 46, timed, w_gready[2]);
                ++ev[3];
              }
            }
            tc_fence_after();
            constexpr uint32_t idesc5 = umma_idesc_bf16(256, F5_N);
            const uint32_t d = tmem_base + (uint32_t)(acc * NB);
            const uint32_t hi_b5 = desc_hi(128, 0);
            const uint32_t w5lo = desc_lo(w_base + wconv, 128);        // one 8-row group per CTA: K core matrices 128 bytes apart
            const uint32_t xa = xlo_base + (uint32_t)(((q_base + r - rho0) % NXR) * (xrow_bytes >> 4));
#pragma unroll
            for (int ks = 0; ks < nx; ++ks)
              umma2_bf16_elect(d, desc_join(xa + (uint32_t)(ks * (SLAB_ROW >> 4)), hi_a), desc_join(w5lo + (uint32_t)(ks * (F5_TILE >> 4)), hi_b5),
                               idesc5, ks > 0 ? 1u : 0u);
#pragma unroll
            for (int g = 0; g < 4; ++g) {
              const int lo_g = r0 - (L - 1 - g) > 0 ? r0 - (L - 1 - g) : 0;
              const uint32_t ga = tmem_base + (uint32_t)(ringcol_of(SCH, g) + ((cb[g] + (r - lo_g)) % ring_of(SCH, g)) * 16);
#pragma unroll
              for (int half = 0; half < 2; ++half)
                umma2_ts_bf16_elect(d, ga + (uint32_t)(half * 8), desc_join(w5lo + (uint32_t)((nx + 2 * g + half) * (F5_TILE >> 4)), hi_b5), idesc5, 1u);
            }
            umma2_commit_elect(tfull(acc));
            ++gcnt;
          };
          if constexpr (F5) {
            group(IC<2>{});
            group(IC<0>{});
            group5();
            group(IC<1>{});
            group(IC<3>{});
          } else {
            group(IC<order_of(SCH, 0)>{});
            group(IC<order_of(SCH, 1)>{});
            group(IC<order_of(SCH, 2)>{});
            if constexpr (L > 3) group(IC<order_of(SCH, 3)>{});
          }
          // X row s - XOLD has no reader left (the next step's oldest row is s + 1 - XOLD)
          const int f = s - XOLD;
          if (f >= rho0 && f < rho1) umma2_commit_elect(xempty((q_base + f - rho0) % NXR));
        }
        for (int f = (s_last - XOLD + 1 > rho0 ? s_last - XOLD + 1 : rho0); f < rho1; ++f) umma2_commit_elect(xempty((q_base + f - rho0) % NXR));
        q_base += rho1 - rho0;
      }
      if (timed && lane == 0) {
        p.dbg[2] = w_tempty; p.dbg[3] = w_xfull; p.dbg[4] = w_gready[0]; p.dbg[5] = w_gready[1]; p.dbg[6] = w_gready[2];
        p.dbg[7] = clock64() - t_start; p.dbg[8] = nsteps; p.dbg[9] = gcnt;
      }
    }
  } else {
    // ===================== epilogue warps 2..17: two teams x two channel halves x four lane quarters =====================
    // Team t drains the groups with (running group count & 1) == t, i.e. accumulator t: two rows' epilogues are in flight at a time.
    const int ew = warp - 2;
    const int team = ew >> 3;
    const int wg = (ew >> 2) & 1;                // output channels 16 wg .. 16 wg + 15
    const int q = warp & 3;                      // TMEM lane quarter = positions 32 q .. 32 q + 31 of the strip row
    const int i = q * 32 + lane;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16);
    const uint32_t tempty_leader = mapa_u32(tempty(team), 0);
    const uint32_t gready_leader0 = mapa_u32(gready(0), 0);
    const uint32_t tfull_bar = tfull(team);
    const uint32_t trow = lane_addr + (uint32_t)(team * NB + wg * 16);
    const int bar_id = 1 + team * 2 + wg;
    const size_t slab_elems = (size_t)p.slabM * 16;
    const bool edge_l = lane == 0, edge_r = lane == 31;
    const bool has_l = q > 0, has_r = q < 3;     // the strip's first / last position has no neighbour
    pdl_wait();
    int gcnt = 0;
    int cnt[MAXL] = {0, 0, 0, 0};         // rows of layer j stored so far: ring slot = count % ring (the MMA warp counts the same)
    int cur = pr_begin;
    Seg sg;
    long long w_tfull = 0;
    int mine = 0, xrows = 0;              // groups / exchanging rows this team has drained
    const long long t_start = DBG ? clock64() : 0;
    while (next_seg(cur, pr_end, h, crank, sg)) {
      const int r0 = sg.r0, r1 = sg.r1;
      const bool col_ok = sg.col < p.ncol;
      const int n = col_ok ? sg.col / p.S : 0;
      const int x = (col_ok ? (sg.col % p.S) : 0) * VALID - HALO + i;
      const bool inimg = col_ok && x >= 0 && x < p.w;
      const bool store_col = inimg && i >= HALO && i < MPOS - HALO;
      const uint32_t keep = inimg ? 0xffffffffu : 0u;       // columns outside the image are the next layer's zero padding
      __nv_bfloat16* ocol = obuf + (size_t)(nx + wg) * slab_elems + ((size_t)n * h * p.w + x) * 16;
      const int s_first = r0 - (L - 1) > 0 ? r0 - (L - 1) : 0, s_last = r1 - 1 + DMAX;
      for (int s = s_first; s <= s_last; ++s) {
        auto group = [&](auto JC) {
          constexpr int J = decltype(JC)::value;
          constexpr bool TO_RING = J < L - 1 || F5;                 // this layer's rows are A operands of later groups
          const int r = s - lag_of(SCH, J);
          const int lo = r0 - (L - 1 - J) > 0 ? r0 - (L - 1 - J) : 0, hi = r1 + (L - 1 - J) < h ? r1 + (L - 1 - J) : h;
          if (r < lo || r >= hi) return;
          const int g = gcnt++;
          int slot = 0;
          if constexpr (TO_RING) slot = cnt[J]++ % ring_of(SCH, J);
          if ((g & 1) != team) return;
          const uint32_t use = (uint32_t)(g >> 1);
          mbar_wait_t(tfull_bar, use & 1u, p.err, 47, timed, w_tfull);
          tc_fence_after();
          uint32_t a0[16], a1[16], a2[16];
          tmem_ld16(trow, a0);
          tmem_ld16(trow + NOUT, a1);
          tmem_ld16(trow + 2 * NOUT, a2);
          tmem_ld_wait();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_cluster(tempty_leader);
          float v[16];
          if constexpr (xs_of(SCH)) {
            // every lane's kx = 0 / kx = 2 partial sums through shared memory: thread i reads L of i - 1 and R of i + 1
            float4* lpl = reinterpret_cast<float4*>(xsbuf + (size_t)(team * 2 + wg) * (XS_BYTES_PER / 4));
            float4* rpl = lpl + (MPOS + 1) * (XS_ROW / 4);
            named_bar_sync(bar_id, 128);             // the previous row's values have been read by everyone
#pragma unroll
            for (int t = 0; t < 16; t += 4) {
              lpl[(i + 1) * (XS_ROW / 4) + t / 4] = make_float4(__uint_as_float(a0[t]), __uint_as_float(a0[t + 1]), __uint_as_float(a0[t + 2]), __uint_as_float(a0[t + 3]));
              rpl[i * (XS_ROW / 4) + t / 4] = make_float4(__uint_as_float(a2[t]), __uint_as_float(a2[t + 1]), __uint_as_float(a2[t + 2]), __uint_as_float(a2[t + 3]));
            }
            named_bar_sync(bar_id, 128);
#pragma unroll
            for (int t = 0; t < 16; t += 4) {
              const float4 l = lpl[i * (XS_ROW / 4) + t / 4], rr = rpl[(i + 1) * (XS_ROW / 4) + t / 4];
              const float4 b = *reinterpret_cast<const float4*>(sbias + J * NOUT + wg * 16 + t);
              v[t] = lrelu02(l.x + __uint_as_float(a1[t]) + rr.x + b.x);
              v[t + 1] = lrelu02(l.y + __uint_as_float(a1[t + 1]) + rr.y + b.y);
              v[t + 2] = lrelu02(l.z + __uint_as_float(a1[t + 2]) + rr.z + b.z);
              v[t + 3] = lrelu02(l.w + __uint_as_float(a1[t + 3]) + rr.w + b.w);
            }
          } else {
          // neighbours across the quarter boundaries: lane 31's kx=0 partials go right, lane 0's kx=2 partials go left.
          // Branch-free on purpose: a divergent region costs ~70 cycles here and this is the per-row critical path.
          // double buffered by the count of rows THIS team exchanged (not by the group count: the conv5-taps group of SCH 3 takes
          // part in the group sequence without a barrier, so two exchanging rows of a team can be 2 groups apart with equal parity)
          float* xb = xbuf + (size_t)(((team * 2 + wg) * 2 + (xrows & 1)) * 4) * 32;
          if (edge_l || edge_r) {
            float4* dst = reinterpret_cast<float4*>(xb + (q * 2 + (edge_l ? 1 : 0)) * 16);
#pragma unroll
            for (int t = 0; t < 16; t += 4)
              dst[t / 4] = edge_l ? make_float4(__uint_as_float(a2[t]), __uint_as_float(a2[t + 1]), __uint_as_float(a2[t + 2]), __uint_as_float(a2[t + 3]))
                                  : make_float4(__uint_as_float(a0[t]), __uint_as_float(a0[t + 1]), __uint_as_float(a0[t + 2]), __uint_as_float(a0[t + 3]));
          }
          named_bar_sync(bar_id, 128);
          // every lane reads both edge vectors of its quarter's neighbours (broadcast loads), lanes 0 / 31 select them
          float el[16], er[16];
          {
            const float4* pl = reinterpret_cast<const float4*>(xb + ((has_l ? q - 1 : 0) * 2 + 0) * 16);
            const float4* pr = reinterpret_cast<const float4*>(xb + ((has_r ? q + 1 : 3) * 2 + 1) * 16);
#pragma unroll
            for (int t = 0; t < 16; t += 4) {
              const float4 a = pl[t / 4], b = pr[t / 4];
              el[t] = a.x; el[t + 1] = a.y; el[t + 2] = a.z; el[t + 3] = a.w;
              er[t] = b.x; er[t + 1] = b.y; er[t + 2] = b.z; er[t + 3] = b.w;
            }
          }
#pragma unroll
          for (int t = 0; t < 16; ++t) {
            const float sl = __shfl_up_sync(0xffffffffu, __uint_as_float(a0[t]), 1);
            const float sr = __shfl_down_sync(0xffffffffu, __uint_as_float(a2[t]), 1);
            const float left = edge_l ? (has_l ? el[t] : 0.f) : sl;
            const float right = edge_r ? (has_r ? er[t] : 0.f) : sr;
            v[t] = lrelu02(left + __uint_as_float(a1[t]) + right + sbias[J * NOUT + wg * 16 + t]);
          }
          }
          uint32_t pk[8];
#pragma unroll
          for (int t = 0; t < 8; ++t) pk[t] = pack_bf2(v[2 * t], v[2 * t + 1]) & keep;
          if constexpr (TO_RING) {
            tmem_st8(lane_addr + (uint32_t)(ringcol_of(SCH, J) + slot * 16 + wg * 8), pk);
            tmem_st_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(gready_leader0 + 8u * (uint32_t)(ringslot0_of(SCH, J) + slot));
          }
          if (store_col && r >= r0 && r < r1 && (!F5 || p.store_growth)) {
            __nv_bfloat16* o = ocol + (size_t)(2 * J) * slab_elems + (size_t)r * p.w * 16;
            *reinterpret_cast<uint4*>(o) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
            *reinterpret_cast<uint4*>(o + 8) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
          }
          ++mine;
          ++xrows;
        };
        // SCH 3: the conv5-taps group -- 9 partial products per pixel (columns tap * 3 + co) -> part[tap][m] as (co0, co1, co2, 0)
        auto group5 = [&]() {
          const int r = s - lag_of(SCH, 4);
          if (r < r0 || r >= r1) return;
          const int g = gcnt++;
          if ((g & 1) != team) return;
          const uint32_t use = (uint32_t)(g >> 1);
          mbar_wait_t(tfull_bar, use & 1u, p.err, 47, timed, w_tfull);
          tc_fence_after();
          uint32_t a[16];
          if (wg == 0) {                       // warp-uniform: the second channel half has nothing to read, it only releases
            tmem_ld16(lane_addr + (uint32_t)(team * NB), a);
            tmem_ld_wait();
          }
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_cluster(tempty_leader);
          if (wg == 0 && store_col) {
            const size_t Mtot = (size_t)p.slabM;
            float* o = p.part + (((size_t)n * h + r) * p.w + x) * 4;
            store4(o, make_float4(__uint_as_float(a[0]), __uint_as_float(a[1]), __uint_as_float(a[2]), 0.f));
            store4(o + Mtot * 4, make_float4(__uint_as_float(a[3]), __uint_as_float(a[4]), __uint_as_float(a[5]), 0.f));
            store4(o + Mtot * 8, make_float4(__uint_as_float(a[6]), __uint_as_float(a[7]), __uint_as_float(a[8]), 0.f));
          }
          ++mine;
        };
        if constexpr (F5) {
          group(IC<2>{});
          group(IC<0>{});
          group5();
          group(IC<1>{});
          group(IC<3>{});
        } else {
          group(IC<order_of(SCH, 0)>{});
          group(IC<order_of(SCH, 1)>{});
          group(IC<order_of(SCH, 2)>{});
          if constexpr (L > 3) group(IC<order_of(SCH, 3)>{});
        }
      }
    }
    if (timed && (threadIdx.x == 64 || threadIdx.x == 64 + 256)) {
      long long* d = p.dbg + 10 + 3 * team;
      d[0] = w_tfull; d[1] = clock64() - t_start; d[2] = mine;
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

// conv5 weight [3][cin_ref = 176][3 taps] fp32 -> bf16 B image of the conv5-taps GEMM, rows n' = tap * 3 + co (9 of 16), split in two
// halves of 8 rows for the CTA pair: [half(2)][kstep(11)][kcore(2)][row % 8][k % 8]
__global__ void pack_f5_kernel(const float* __restrict__ wref, __nv_bfloat16* __restrict__ img, int cin) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= F5_N * cin) return;
  const int n = idx % F5_N, c = idx / F5_N;
  const int tap = n / 3, co = n % 3;
  const float v = n < 9 ? wref[((size_t)co * cin + c) * 3 + tap] : 0.f;
  const int ks = c / 16, kk = c % 16;
  img[(size_t)(n / 8) * (F5_KSTEPS * 128) + (size_t)ks * 128 + (kk / 8) * 64 + (n % 8) * 8 + (kk % 8)] = __float2bfloat16_rn(v);
}

// F(x2)[t] = sum_tap W[tap] . in[t + tap - 1] from the per-frame partial products, then the additive half of the coupling
// (SelfC_GMM_arch_inv.py:24,31): y1 = x1 +/- (F + bias) on the fp32 latent state, with bf16 copies of y1 into the X slabs of the
// coupling's G and H dense buffers (channels 3..15 zero) -- what EPI_COUPLE_Y1 of the temporal kernel does for the layer-by-layer path
__global__ void __launch_bounds__(256) f5_combine_kernel(const float* __restrict__ part, const float* __restrict__ bias, float* __restrict__ z,
                                                         __nv_bfloat16* __restrict__ copyA, __nv_bfloat16* __restrict__ copyB, int T, long long hw,
                                                         long long M, int rev) {
  const long long m = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= M) return;
  const int t = (int)((m / hw) % T);
  float4 f = *reinterpret_cast<const float4*>(part + ((size_t)M + m) * 4);                    // tap 1: this frame
  if (t > 0) {
    const float4 q = *reinterpret_cast<const float4*>(part + (size_t)(m - hw) * 4);           // tap 0 applied to frame t - 1
    f.x += q.x; f.y += q.y; f.z += q.z;
  }
  if (t + 1 < T) {
    const float4 q = *reinterpret_cast<const float4*>(part + ((size_t)2 * M + m + hw) * 4);    // tap 2 applied to frame t + 1
    f.x += q.x; f.y += q.y; f.z += q.z;
  }
  f.x += __ldg(bias); f.y += __ldg(bias + 1); f.z += __ldg(bias + 2);
  float* zp = z + quad_off((size_t)M, 0, (size_t)m);
  const float4 x1 = *reinterpret_cast<const float4*>(zp);
  const float y0 = rev ? x1.x - f.x : x1.x + f.x, y1 = rev ? x1.y - f.y : x1.y + f.y, y2 = rev ? x1.z - f.z : x1.z + f.z;
  *reinterpret_cast<float4*>(zp) = make_float4(y0, y1, y2, 0.f);
  const uint4 lo = make_uint4(pack_bf2(y0, y1), pack_bf2(y2, 0.f), 0u, 0u), zero = make_uint4(0u, 0u, 0u, 0u);
  if (copyA) {
    uint4* o = reinterpret_cast<uint4*>(copyA + (size_t)m * 16);
    o[0] = lo;
    o[1] = zero;
  }
  if (copyB) {
    uint4* o = reinterpret_cast<uint4*>(copyB + (size_t)m * 16);
    o[0] = lo;
    o[1] = zero;
  }
}

static int smem_bytes(int sch, int nx) {
  const int L = nlayers_of(sch);
  return 1024 + nxr_of(sch) * nx * SLAB_ROW + 3 * WT_BYTES * (L * nx + L * (L - 1)) + BAR_BYTES + XBUF_BYTES + BIAS_BYTES +
         (xs_of(sch) ? 4 * XS_BYTES_PER : 0) + (f5_of(sch) ? F5_KSTEPS * F5_TILE : 0);
}
// One instantiation per X width the network has -- 1 slab (G, H, local_m1), 3 (F), 4 (the STP 64 -> 64 blocks) -- each with the
// deepest schedule whose X ring and weights fit shared memory; -1: not built for this width (the caller runs layer by layer).
static int pick_schedule(int nx, int max_layers) {
  const int sch = nx == 1 ? 0 : nx == 3 ? 1 : nx == 4 ? 2 : -1;
  if (sch < 0 || nlayers_of(sch) > max_layers || smem_bytes(sch, nx) > 227 * 1024) return -1;
  return sch;
}

}  // namespace dbf

int dense_fused_schedule(int sch, int* out) {
  if (sch < 0 || sch > 3 || out == nullptr) return SELFC_E_ARG;
  const int L = dbf::nlayers_of(sch), ng = dbf::ngroups_of(sch);
  for (int i = 0; i < 18; ++i) out[i] = 0;
  out[0] = L;
  out[1] = ng;
  out[2] = dbf::nxr_of(sch);
  for (int j = 0; j < ng; ++j) out[3 + j] = dbf::lag_of(sch, j);
  for (int oi = 0; oi < ng; ++oi) out[8 + oi] = dbf::order_of(sch, oi);
  const int kept = dbf::f5_of(sch) ? L : L - 1;          // layers whose rows are A operands of later groups
  int cols = dbf::RING_COL0;
  for (int j = 0; j < kept; ++j) {
    out[13 + j] = dbf::ring_of(sch, j);
    cols += 16 * dbf::ring_of(sch, j);
  }
  out[17] = cols;
  return 0;
}

int dense_fused_layers(int cin) {
  if (cin % 16 != 0 || cin <= 0) return 0;
  const int sch = dbf::pick_schedule(cin / 16, 4);
  return sch < 0 ? 0 : dbf::nlayers_of(sch);
}

int pack_f5_weights(void** img, const float* wref, int cin, cudaStream_t st) {
  SELFC_CHECK_ARG(cin == 16 * dbf::F5_KSTEPS, "pack_f5_weights: built for %d input channels, got %d", 16 * dbf::F5_KSTEPS, cin);
  const size_t bytes = (size_t)2 * dbf::F5_KSTEPS * dbf::F5_TILE;
  if (*img == nullptr) SELFC_CUDA(cudaMalloc(img, bytes));
  dbf::pack_f5_kernel<<<cdiv(dbf::F5_N * cin, 256), 256, 0, st>>>(wref, reinterpret_cast<__nv_bfloat16*>(*img), cin);
  SELFC_LAUNCH_CHECK("pack_f5_kernel");
  return 0;
}

int launch_f5_combine(const float* part, const float* bias, float* z, __nv_bfloat16* copyA, __nv_bfloat16* copyB, int T, long long hw,
                      long long M, int rev, cudaStream_t st) {
  if (M == 0) return 0;
  SELFC_CHECK_ARG(part && bias && z && aligned16(part) && aligned16(z), "f5_combine: null or misaligned buffer");
  dbf::f5_combine_kernel<<<cdiv(M, 256), 256, 0, st>>>(part, bias, z, copyA, copyB, T, hw, M, rev);
  SELFC_LAUNCH_CHECK("f5_combine_kernel");
  return 0;
}

int launch_dense_fused(const TcConvW* w, int L, __nv_bfloat16* buf, long long slabM, int cin, int N, int h, int wd, cudaStream_t st,
                       const TcConvW* w2, __nv_bfloat16* buf2, const void* f5img, float* f5part) {
  SELFC_CHECK_ARG(cin % 16 == 0 && cin > 0, "dense_fused: cin=%d must be a positive multiple of 16", cin);
  const int nx = cin / 16;
  const bool f5 = f5img != nullptr;
  SELFC_CHECK_ARG(!f5 || (nx == 3 && L == 4 && w2 == nullptr && f5part != nullptr && aligned16(f5part)),
                  "dense_fused: the conv5-taps group is built for the F block (cin 48, 4 layers, single problem)");
  const int sch = f5 ? 3 : dbf::pick_schedule(nx, L);
  SELFC_CHECK_ARG(!f5 || dbf::smem_bytes(3, nx) <= 227 * 1024, "dense_fused: the conv5-taps schedule does not fit shared memory");
  SELFC_CHECK_ARG(sch >= 0 && dbf::nlayers_of(sch) == L, "dense_fused: cin=%d with %d fused layers does not fit shared memory", cin, L);
  SELFC_CHECK_ARG(aligned16(buf) && slabM == (long long)N * h * wd, "dense_fused: slab layout / alignment");
  const bool dual = w2 != nullptr;
  SELFC_CHECK_ARG(!dual || (buf2 != nullptr && aligned16(buf2) && buf2 != buf), "dense_fused: the second problem needs its own buffer");
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
  p.w5img = f5img;
  p.part = f5part;
  p.store_growth = f5 ? 0 : 1;
  p.err = tc::err_flag_for_device();
  const bool dbg = tc::debug_slots();
  if (dbg) p.dbg = tc::debug_next_slot(8000000 + (dual ? 100000 : 0) + sch * 1000 + nx);
  const int smem = dbf::smem_bytes(sch, nx);
  static bool smem_set[64] = {};           // per device: function attributes belong to the device's context
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev >= 0 && dev < 64 && !smem_set[dev]) {
    SELFC_CUDA(cudaFuncSetAttribute(dbf::dense_fused_kernel<0, 1, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    SELFC_CUDA(cudaFuncSetAttribute(dbf::dense_fused_kernel<1, 3, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    SELFC_CUDA(cudaFuncSetAttribute(dbf::dense_fused_kernel<2, 4, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    SELFC_CUDA(cudaFuncSetAttribute(dbf::dense_fused_kernel<0, 1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    SELFC_CUDA(cudaFuncSetAttribute(dbf::dense_fused_kernel<1, 3, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    SELFC_CUDA(cudaFuncSetAttribute(dbf::dense_fused_kernel<2, 4, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    SELFC_CUDA(cudaFuncSetAttribute(dbf::dense_fused_kernel<3, 3, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    SELFC_CUDA(cudaFuncSetAttribute(dbf::dense_fused_kernel<3, 3, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    smem_set[dev] = true;
  }
  const int grid = 2 * pieces * p.nprob;
  auto kern = sch == 0 ? (dbg ? dbf::dense_fused_kernel<0, 1, true> : dbf::dense_fused_kernel<0, 1, false>)
            : sch == 1 ? (dbg ? dbf::dense_fused_kernel<1, 3, true> : dbf::dense_fused_kernel<1, 3, false>)
            : sch == 3 ? (dbg ? dbf::dense_fused_kernel<3, 3, true> : dbf::dense_fused_kernel<3, 3, false>)
                       : (dbg ? dbf::dense_fused_kernel<2, 4, true> : dbf::dense_fused_kernel<2, 4, false>);
  SELFC_CUDA(tc::launch_pdl_pairs(kern, grid, dbf::THREADS, smem, st, tmap, tmap2, p));
  SELFC_LAUNCH_CHECK("dense_fused_kernel");
  return 0;
}

}  // namespace selfc
