// fp32-FMA implicit-GEMM convolution over pixel-major buffers (the strict-fp32 numerics mode, and the small
// GEMMs -- GlobalAgg proj1 mix -- of both modes).  One kernel covers the four tap patterns of the path:
//   TAP_SPATIAL  (1,3,3) convs conv1..4 of D2DTInput            Subnet_constructor.py:102-105,126-129
//   TAP_TEMPORAL (3,1,1) conv5 of D2DTInput, zero-padded in t    Subnet_constructor.py:106,130
//   TAP_POINT    1x1x1 convs of the GMM head                     SelfC_GMM_arch_inv.py:336-344
//   TAP_TMIX     GlobalAgg: proj1 applied to the T-mixed input   SelfC_GMM_arch_inv.py:266,278-285
// and fuses the coupling arithmetic of InvBlockExp (SelfC_GMM_arch_inv.py:21-33) into the conv5 epilogue.
//
// Tile: 128 pixels x 32 output channels per CTA, 256 threads, 4x4 register tile, K staged through shared
// memory in chunks of 16 with register prefetch (global loads of chunk i+1 overlap the FMAs of chunk i).
#include "common.cuh"
#include "kernels.h"
#include "pack_batch.h"

namespace selfc {

constexpr int SM_TILE_M = 128;
constexpr int SM_TILE_N = 32;
constexpr int SM_KC = 16;
constexpr int SM_APITCH = SM_TILE_M + 4;

template <typename T>
__global__ void __launch_bounds__(256, 3) conv_simt_kernel(const ConvArgs<T> a) {
  __shared__ __align__(16) float As[2][SM_KC][SM_APITCH];
  __shared__ __align__(16) float Bs[2][SM_KC][SM_TILE_N];

  const int tid = threadIdx.x;
  const int tx = tid & 7;    // 4 output channels
  const int ty = tid >> 3;   // 4 pixels
  const long long hw = (long long)a.h * a.w_;
  const long long M = (long long)a.BT * hw;
  const long long m0 = (long long)blockIdx.x * SM_TILE_M;
  const int nb = blockIdx.y * SM_TILE_N;
  const int ktot = a.taps * a.cin;
  const int nchunk = (ktot + SM_KC - 1) / SM_KC;

  // loader role: this thread stages pixel lp, k-groups g0 and g0+2 of every chunk
  const int lp = tid & 127;
  const int g0 = tid >> 7;
  const long long lm = m0 + lp;
  const bool lvalid = lm < M;
  int ly = 0, lx = 0, lt = 0, lb = 0;
  if (lvalid) {
    long long n = lm / hw;
    long long pix = lm - n * hw;
    ly = (int)(pix / a.w_);
    lx = (int)(pix - (long long)ly * a.w_);
    lt = (int)(n % a.Tn);
    lb = (int)(n / a.Tn);
  }
  // B loader role
  const int bk = tid >> 4;
  const int bn = (tid & 15) * 2;

  float4 ra[2];
  float2 rb;

  auto load_chunk = [&](int ck) {
    const int k0 = ck * SM_KC;
#pragma unroll
    for (int gi = 0; gi < 2; ++gi) {
      const int k = k0 + (g0 + 2 * gi) * 4;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (lvalid && k < ktot) {
        const int tap = k / a.cin;
        const int c = k - tap * a.cin;
        if (a.tap_mode == TAP_TMIX) {
          const float* wm = a.wmat + ((long long)lb * a.Tn) * a.Tn + lt;   // W[b][tt][t'], t' = lt
          for (int tt = 0; tt < a.Tn; ++tt) {
            const float wv = __ldg(wm + tt * a.Tn);
            float4 r = load4(a.in + dense_off(lm + (long long)(tt - lt) * hw, c, a.in_pitch, a.in_slabM));
            v.x += wv * r.x; v.y += wv * r.y; v.z += wv * r.z; v.w += wv * r.w;
          }
        } else {
          bool ok = true;
          long long src = lm;
          if (a.tap_mode == TAP_SPATIAL) {
            const int dy = tap / 3 - 1, dx = tap % 3 - 1;
            ok = (unsigned)(ly + dy) < (unsigned)a.h && (unsigned)(lx + dx) < (unsigned)a.w_;
            src = lm + dy * a.w_ + dx;
          } else if (a.tap_mode == TAP_TEMPORAL) {
            const int dt = tap - 1;
            ok = (unsigned)(lt + dt) < (unsigned)a.Tn;
            src = lm + (long long)dt * hw;
          }
          if (ok) v = load4(a.in + dense_off(src, c, a.in_pitch, a.in_slabM));
        }
        if (a.in_lrelu) { v.x = lrelu02(v.x); v.y = lrelu02(v.y); v.z = lrelu02(v.z); v.w = lrelu02(v.w); }
      }
      ra[gi] = v;
    }
    const int kb = k0 + bk;
    rb = make_float2(0.f, 0.f);
    if (kb < ktot) rb = __ldg(reinterpret_cast<const float2*>(a.w + (long long)kb * a.np + nb + bn));
  };
  auto store_chunk = [&](int buf) {
#pragma unroll
    for (int gi = 0; gi < 2; ++gi) {
      const int kk = (g0 + 2 * gi) * 4;
      As[buf][kk + 0][lp] = ra[gi].x;
      As[buf][kk + 1][lp] = ra[gi].y;
      As[buf][kk + 2][lp] = ra[gi].z;
      As[buf][kk + 3][lp] = ra[gi].w;
    }
    *reinterpret_cast<float2*>(&Bs[buf][bk][bn]) = rb;
  };

  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  load_chunk(0);
  store_chunk(0);
  __syncthreads();
  for (int ck = 0; ck < nchunk; ++ck) {
    const int buf = ck & 1;
    if (ck + 1 < nchunk) load_chunk(ck + 1);
#pragma unroll
    for (int kk = 0; kk < SM_KC; ++kk) {
      const float4 av = *reinterpret_cast<const float4*>(&As[buf][kk][ty * 4]);
      const float4 bv = *reinterpret_cast<const float4*>(&Bs[buf][kk][tx * 4]);
      const float ar[4] = {av.x, av.y, av.z, av.w};
      const float br[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(ar[i], br[j], acc[i][j]);
    }
    if (ck + 1 < nchunk) store_chunk(buf ^ 1);
    __syncthreads();
  }

  // ---- epilogue ------------------------------------------------------------------------------------------
  const int n0 = nb + tx * 4;
  const float4 bias4 = __ldg(reinterpret_cast<const float4*>(a.bias + n0));
  const float bias[4] = {bias4.x, bias4.y, bias4.z, bias4.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const long long m = m0 + ty * 4 + i;
    if (m >= M) continue;
    float v[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) v[j] = acc[i][j] + bias[j];
    switch (a.epi) {
      case EPI_STORE: {
        if (a.act) {
#pragma unroll
          for (int j = 0; j < 4; ++j) v[j] = lrelu02(v[j]);
        }
        if (a.outT) {
          T* o = a.outT + dense_off(m, a.outT_off + n0, a.outT_pitch, a.outT_slabM);
          if (n0 + 4 <= a.cout) store4(o, make_float4(v[0], v[1], v[2], v[3]));
          else
            for (int j = 0; j < 4; ++j)
              if (n0 + j < a.cout) o[j] = from_f<T>(v[j]);
        }
        if (a.outF) {
          float* o = a.outF + m * a.outF_pitch + a.outF_off + n0;
          if (n0 + 4 <= a.cout) store4(o, make_float4(v[0], v[1], v[2], v[3]));
          else
            for (int j = 0; j < 4; ++j)
              if (n0 + j < a.cout) o[j] = v[j];
        }
      } break;
      case EPI_ACCUM: {       // outF[m][off + n] += v   (dgrad of the training step: gradients of a concat sum up)
        float* o = a.outF + m * a.outF_pitch + a.outF_off + n0;
        for (int j = 0; j < 4; ++j)
          if (n0 + j < a.cout) o[j] += v[j];
      } break;
      case EPI_COUPLE_Y1: {   // y1 = x1 +/- F(x2)   (SelfC_GMM_arch_inv.py:25, :31)
        if (n0 < a.copy_pad || n0 < 4) {
          float y[4] = {0.f, 0.f, 0.f, 0.f};
          if (n0 == 0) {
            float* zp = a.z + quad_off(M, 0, m);
            const float4 x1 = load4(zp);
            y[0] = a.rev ? x1.x - v[0] : x1.x + v[0];
            y[1] = a.rev ? x1.y - v[1] : x1.y + v[1];
            y[2] = a.rev ? x1.z - v[2] : x1.z + v[2];
            store4(zp, make_float4(y[0], y[1], y[2], 0.f));
          }
          if (n0 < a.copy_pad) {
            const float4 yv = make_float4(y[0], y[1], y[2], 0.f);
            if (a.copyA) store4(a.copyA + dense_off(m, n0, a.copyA_pitch, a.copy_slabM), yv);
            if (a.copyB) store4(a.copyB + dense_off(m, n0, a.copyB_pitch, a.copy_slabM), yv);
          }
        }
      } break;
      case EPI_COUPLE_S: {    // s = clamp * (2*sigmoid(H) - 1), clamp = 1   (:26, :29)
        if (n0 < kHF) {
          float s[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) s[j] = (1.0f / (1.0f + expf(-v[j]))) * 2.0f - 1.0f;
          store4(a.sbuf + quad_off(M, n0 / 4, m), make_float4(s[0], s[1], s[2], s[3]));
        }
      } break;
      case EPI_COUPLE_Y2: {   // y2 = x2*exp(s) + G(y1)  |  y2 = (x2 - G(x1)) / exp(s)   (:27, :30)
        if (n0 < kHF) {
          float* zp = a.z + quad_off(M, 1 + n0 / 4, m);
          const float4 x2 = load4(zp);
          const float4 s4 = load4(a.sbuf + quad_off(M, n0 / 4, m));
          const float xr[4] = {x2.x, x2.y, x2.z, x2.w};
          const float sr[4] = {s4.x, s4.y, s4.z, s4.w};
          float y[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float e = expf(sr[j]);
            y[j] = a.rev ? (xr[j] - v[j]) / e : xr[j] * e + v[j];
          }
          const float4 yv = make_float4(y[0], y[1], y[2], y[3]);
          store4(zp, yv);
          if (a.copyA) store4(a.copyA + dense_off(m, n0, a.copyA_pitch, a.copy_slabM), yv);
        }
      } break;
      case EPI_GA: {          // out = x + proj1(mix) + bias * colsum(W)   (:266, :278-285 by linearity)
        if (n0 < a.cout) {
          const long long n = m / hw;
          const float ws = __ldg(a.wsum + n);   // [B][T] flattened == frame index
          const float4 r = load4(a.resid + m * a.resid_pitch + n0);
          const float rr[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
          for (int j = 0; j < 4; ++j) v[j] = rr[j] + acc[i][j] + bias[j] * ws;
          const float4 ov = make_float4(v[0], v[1], v[2], v[3]);
          if (a.outT) store4(a.outT + dense_off(m, a.outT_off + n0, a.outT_pitch, a.outT_slabM), ov);
          if (a.outF) store4(a.outF + m * a.outF_pitch + a.outF_off + n0, ov);
          if (a.outAct)
            store4(a.outAct + m * a.outAct_pitch + n0, make_float4(lrelu02(v[0]), lrelu02(v[1]), lrelu02(v[2]), lrelu02(v[3])));
        }
      } break;
    }
  }
}

template <typename T>
int launch_conv_simt(const ConvArgs<T>& a, cudaStream_t st) {
  SELFC_CHECK_ARG(a.cin % 4 == 0 && a.in_pitch % 4 == 0 && a.np % SM_TILE_N == 0, "conv_simt: cin/pitch/np alignment");
  const long long M = (long long)a.BT * a.h * a.w_;
  if (M == 0) return 0;
  dim3 grid(cdiv(M, SM_TILE_M), a.np / SM_TILE_N);
  if (a.epi != EPI_STORE && a.epi != EPI_GA && a.epi != EPI_ACCUM) {
    // coupling epilogues only consume the first channels: skip all-padding column tiles
    int need = (a.epi == EPI_COUPLE_Y1) ? (a.copy_pad > 4 ? a.copy_pad : 4) : kHF;
    grid.y = cdiv(need, SM_TILE_N);
  }
  conv_simt_kernel<T><<<grid, 256, 0, st>>>(a);
  SELFC_LAUNCH_CHECK("conv_simt_kernel");
  return 0;
}
template int launch_conv_simt<float>(const ConvArgs<float>&, cudaStream_t);
template int launch_conv_simt<__nv_bfloat16>(const ConvArgs<__nv_bfloat16>&, cudaStream_t);
// BF16X3 mode runs on the tcgen05 kernels only: there is no FMA fallback for (hi, lo) buffers, and asking for one is an error
template <>
int launch_conv_simt<bfx2>(const ConvArgs<bfx2>&, cudaStream_t) {
  set_error("BF16X3 mode: this convolution has no tcgen05 path for the requested shape (no fp32-FMA fallback for (hi, lo) buffers)");
  return SELFC_E_UNSUPPORTED;
}

// ---- weight packing ---------------------------------------------------------------------------------------
struct PackSimtJob {
  const float* wref;
  const float* bref;
  float* wpk;
  float* bpk;
  int cout, cin_ref, taps, cin_buf, xreal, xpad, np;
  int pad;                 // (no implicit padding: the job tables are compared bytewise)
};

__device__ __forceinline__ void pack_conv_simt_body(const PackSimtJob& j, long long idx) {
  const long long total = (long long)j.taps * j.cin_buf * j.np;
  if (idx < j.np) j.bpk[idx] = idx < j.cout ? j.bref[idx] : 0.f;
  if (idx >= total) return;
  const int n = (int)(idx % j.np);
  const int c = (int)((idx / j.np) % j.cin_buf);
  const int tap = (int)(idx / ((long long)j.np * j.cin_buf));
  int cref = c < j.xreal ? c : (c < j.xpad ? -1 : c - j.xpad + j.xreal);
  float v = 0.f;
  if (n < j.cout && cref >= 0 && cref < j.cin_ref) v = j.wref[((long long)n * j.cin_ref + cref) * j.taps + tap];
  j.wpk[idx] = v;
}

__global__ void pack_conv_simt_kernel(const PackSimtJob j) { pack_conv_simt_body(j, (long long)blockIdx.x * blockDim.x + threadIdx.x); }

// every recorded job in one launch (pack_batch.h)
__global__ void pack_conv_simt_multi_kernel(const PackSimtJob* __restrict__ jobs, const int* __restrict__ first, int njobs) {
  const int ji = pack_find_job(first, njobs, blockIdx.x);
  const PackSimtJob j = jobs[ji];
  pack_conv_simt_body(j, (long long)(blockIdx.x - __ldg(first + ji)) * blockDim.x + threadIdx.x);
}

int launch_pack_conv_simt(const float* wref, const float* bref, float* wpk, float* bpk, int cout, int cin_ref, int taps,
                          int cin_buf, int xreal, int xpad, int np, cudaStream_t st) {
  const long long total = (long long)taps * cin_buf * np;
  const PackSimtJob j{wref, bref, wpk, bpk, cout, cin_ref, taps, cin_buf, xreal, xpad, np, 0};
  const int nblocks = (int)cdiv(total > np ? total : np, 256);
  if (PackBatch* pb = pack_batch_current()) {
    pb->simt[pb->point].add(j, nblocks);
    return 0;
  }
  pack_conv_simt_kernel<<<nblocks, 256, 0, st>>>(j);
  SELFC_LAUNCH_CHECK("pack_conv_simt_kernel");
  return 0;
}

int flush_pack_conv_simt(JobTable& t, cudaStream_t st) {
  if (t.njobs() <= 0) return 0;
  const void* jobs = nullptr;
  const int* first = nullptr;
  SELFC_CUDA(t.sync(st, &jobs, &first));
  pack_conv_simt_multi_kernel<<<t.first.back(), 256, 0, st>>>(static_cast<const PackSimtJob*>(jobs), first, t.njobs());
  SELFC_LAUNCH_CHECK("pack_conv_simt_multi_kernel");
  return 0;
}

}  // namespace selfc
