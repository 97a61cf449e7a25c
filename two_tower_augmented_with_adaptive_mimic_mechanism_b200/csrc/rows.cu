// Row gather / cast / gate / augmentation / fused loss kernels (HBM-bound, SIMT by design).
//   ttam_gather_rows_f32 : nn.Embedding.forward / index_select     (reference encoders.py:223,
//                          adaptive_mimic.py:97-105, training.py:743-775)
//   ttam_gate_fwd/bwd    : FeatureFusionGate tail + _apply_aug      (encoders.py:160-168, adaptive_mimic.py:88-95)
//   ttam_loss_fwd_bwd    : dot products + BCEWithLogits + 2x MSE    (training.py:770-803, adaptive_mimic.py:66-67)
#include "common.cuh"
#include <initializer_list>

namespace ttam {

// ------------------------------------------------------------------------------------------------
// gather: a warp takes kGatherRows consecutive output rows; their indices are ONE coalesced load by the first lanes
// (no per-chunk division, no per-chunk dependent index load), then lane = 16-byte chunk of the row with all kGatherRows
// row loads issued before the first store (8 x 384 B in flight per warp at D = 96).
// ------------------------------------------------------------------------------------------------
constexpr int kGatherRows = 8;

template <bool VEC>
__global__ void __launch_bounds__(256) gather_rows_kernel(const float* __restrict__ table, int64_t ld_t,
                                                          int64_t num_rows, const int64_t* __restrict__ idx,
                                                          float* __restrict__ out, int64_t ld_o, int64_t R,
                                                          int64_t ncols) {
  if (VEC) {
    const int lane = threadIdx.x & 31;
    const int64_t r0 = (((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5) * kGatherRows;
    if (r0 >= R) return;
    int64_t mine = -1;
    if (lane < kGatherRows && r0 + lane < R) {
      mine = idx[r0 + lane];
      if (mine < 0 || mine >= num_rows) mine = -1;   // rows outside the table come back as zeros
    }
    const int nrows = (int)((R - r0) < kGatherRows ? (R - r0) : kGatherRows);
    const int cpr = (int)(ncols >> 2);  // 16-byte chunks per row
    for (int c0 = 0; c0 < cpr; c0 += 32) {
      const int c = c0 + lane;
      float4 v[kGatherRows];
#pragma unroll
      for (int u = 0; u < kGatherRows; ++u) {
        const int64_t src = __shfl_sync(0xffffffffu, mine, u);
        v[u] = (c < cpr && src >= 0) ? ld_f4(table + src * ld_t + c * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int u = 0; u < kGatherRows; ++u)
        if (c < cpr && u < nrows) st_f4(out + (r0 + u) * ld_o + c * 4, v[u]);
    }
  } else {
    const int64_t total = R * ncols;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (int64_t)gridDim.x * blockDim.x) {
      int64_t r = i / ncols, c = i - r * ncols;
      int64_t src = idx[r];
      out[r * ld_o + c] = (src >= 0 && src < num_rows) ? table[src * ld_t + c] : 0.f;
    }
  }
}

__global__ void __launch_bounds__(256) cast_bf16_kernel(const float* __restrict__ src, int64_t ld_s,
                                                        uint16_t* __restrict__ dst, int64_t ld_d, int64_t R,
                                                        int64_t ncols) {
  const int64_t total = R * ncols;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    int64_t r = i / ncols, c = i - r * ncols;
    uint32_t u = __float_as_uint(src[r * ld_s + c]);
    // round to nearest even; NaN stays NaN
    uint32_t lsb = (u >> 16) & 1u;
    uint32_t rounded = u + 0x7FFFu + lsb;
    uint16_t h = ((u & 0x7F800000u) == 0x7F800000u) ? (uint16_t)((u >> 16) | ((u & 0xFFFFu) ? 0x40u : 0u))
                                                    : (uint16_t)(rounded >> 16);
    dst[r * ld_d + c] = h;
  }
}

// ------------------------------------------------------------------------------------------------
// gate forward / backward, augmentation
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + expf(-x)); }

__global__ void __launch_bounds__(256) gate_fwd_kernel(const float* __restrict__ z, const float* __restrict__ pre2,
                                                       const float* __restrict__ aug, int64_t aug_rows,
                                                       const int64_t* __restrict__ idx, float* __restrict__ g,
                                                       float* __restrict__ t, float* __restrict__ o,
                                                       float* __restrict__ q_out, int64_t R, int64_t D) {
  const int64_t total = R * D;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    int64_t r = i / D, d = i - r * D;
    float e = z[r * 2 * D + d], f = z[r * 2 * D + D + d];
    float gg = sigmoidf_(pre2[i]);
    float tt = gg * e + (1.f - gg) * f;
    g[i] = gg;
    if (t) t[i] = tt;
    float q = 0.f;
    if (aug) {
      int64_t src = idx[r];
      q = (src >= 0 && src < aug_rows) ? aug[src * D + d] : 0.f;
      if (q_out) q_out[i] = q;
    }
    if (o) o[i] = tt + q;
  }
}

__global__ void __launch_bounds__(256) gate_bwd_kernel(const float* __restrict__ dt, const float* __restrict__ z,
                                                       const float* __restrict__ g, float* __restrict__ dpre2,
                                                       float* __restrict__ dz, int64_t R, int64_t D) {
  const int64_t total = R * D;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    int64_t r = i / D, d = i - r * D;
    float e = z[r * 2 * D + d], f = z[r * 2 * D + D + d];
    float gg = g[i], d_t = dt[i];
    dpre2[i] = d_t * (e - f) * gg * (1.f - gg);
    dz[r * 2 * D + d] = d_t * gg;
    dz[r * 2 * D + D + d] = d_t * (1.f - gg);
  }
}

__global__ void __launch_bounds__(256) augment_fwd_kernel(const float* __restrict__ t, const float* __restrict__ aug,
                                                          int64_t aug_rows, const int64_t* __restrict__ idx,
                                                          float* __restrict__ o, float* __restrict__ q_out, int64_t R,
                                                          int64_t D) {
  const int64_t total = R * D;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    int64_t r = i / D, d = i - r * D;
    int64_t src = idx[r];
    float q = (src >= 0 && src < aug_rows) ? aug[src * D + d] : 0.f;
    if (q_out) q_out[i] = q;
    o[i] = t[i] + q;
  }
}

// 16-byte forms of the three kernels above (D % 4 == 0, 16-byte aligned pointers): one float4 chunk per thread-iteration,
// two chunks in flight per thread, no per-element 64-bit division (one per chunk).
__device__ __forceinline__ float4 sigmoid4(const float4& x) {
  return make_float4(sigmoidf_(x.x), sigmoidf_(x.y), sigmoidf_(x.z), sigmoidf_(x.w));
}

__global__ void __launch_bounds__(256) gate_fwd_vec_kernel(const float* __restrict__ z, const float* __restrict__ pre2,
                                                           const float* __restrict__ aug, int64_t aug_rows,
                                                           const int64_t* __restrict__ idx, float* __restrict__ g,
                                                           float* __restrict__ t, float* __restrict__ o,
                                                           float* __restrict__ q_out, int64_t R, int D) {
  const int cpr = D >> 2;
  const int64_t total = R * cpr, stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i0 < total; i0 += 2 * stride) {
    float4 e[2], f[2], p[2], q[2];
    bool ok[2];
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const int64_t i = i0 + u * stride;
      ok[u] = i < total;
      q[u] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (ok[u]) {
        const int64_t r = i / cpr;
        const int c = (int)(i - r * cpr) * 4;
        e[u] = ld_f4(z + r * 2 * D + c);
        f[u] = ld_f4(z + r * 2 * D + D + c);
        p[u] = ld_f4(pre2 + i * 4);
        if (aug) {
          const int64_t src = idx[r];
          if (src >= 0 && src < aug_rows) q[u] = ld_f4(aug + src * D + c);
        }
      }
    }
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      if (!ok[u]) continue;
      const int64_t i = i0 + u * stride;
      const float4 gg = sigmoid4(p[u]);
      const float4 tt = make_float4(gg.x * e[u].x + (1.f - gg.x) * f[u].x, gg.y * e[u].y + (1.f - gg.y) * f[u].y,
                                    gg.z * e[u].z + (1.f - gg.z) * f[u].z, gg.w * e[u].w + (1.f - gg.w) * f[u].w);
      st_f4(g + i * 4, gg);
      if (t) st_f4(t + i * 4, tt);
      if (aug && q_out) st_f4(q_out + i * 4, q[u]);
      if (o) st_f4(o + i * 4, make_float4(tt.x + q[u].x, tt.y + q[u].y, tt.z + q[u].z, tt.w + q[u].w));
    }
  }
}

__global__ void __launch_bounds__(256) gate_bwd_vec_kernel(const float* __restrict__ dt, const float* __restrict__ z,
                                                           const float* __restrict__ g, float* __restrict__ dpre2,
                                                           float* __restrict__ dz, int64_t R, int D) {
  const int cpr = D >> 2;
  const int64_t total = R * cpr, stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i0 < total; i0 += 2 * stride) {
    float4 e[2], f[2], gg[2], d[2];
    int64_t zo[2];
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const int64_t i = i0 + u * stride;
      zo[u] = -1;
      if (i < total) {
        const int64_t r = i / cpr;
        zo[u] = r * 2 * D + (int64_t)(i - r * cpr) * 4;
        e[u] = ld_f4(z + zo[u]);
        f[u] = ld_f4(z + zo[u] + D);
        gg[u] = ld_f4(g + i * 4);
        d[u] = ld_f4(dt + i * 4);
      }
    }
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      if (zo[u] < 0) continue;
      const int64_t i = i0 + u * stride;
      st_f4(dpre2 + i * 4, make_float4(d[u].x * (e[u].x - f[u].x) * gg[u].x * (1.f - gg[u].x),
                                       d[u].y * (e[u].y - f[u].y) * gg[u].y * (1.f - gg[u].y),
                                       d[u].z * (e[u].z - f[u].z) * gg[u].z * (1.f - gg[u].z),
                                       d[u].w * (e[u].w - f[u].w) * gg[u].w * (1.f - gg[u].w)));
      st_f4(dz + zo[u], make_float4(d[u].x * gg[u].x, d[u].y * gg[u].y, d[u].z * gg[u].z, d[u].w * gg[u].w));
      st_f4(dz + zo[u] + D, make_float4(d[u].x * (1.f - gg[u].x), d[u].y * (1.f - gg[u].y), d[u].z * (1.f - gg[u].z),
                                        d[u].w * (1.f - gg[u].w)));
    }
  }
}

__global__ void __launch_bounds__(256) augment_fwd_vec_kernel(const float* __restrict__ t, const float* __restrict__ aug,
                                                              int64_t aug_rows, const int64_t* __restrict__ idx,
                                                              float* __restrict__ o, float* __restrict__ q_out, int64_t R,
                                                              int D) {
  const int cpr = D >> 2;
  const int64_t total = R * cpr, stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i0 < total; i0 += 4 * stride) {
    float4 tt[4], q[4];
    bool ok[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int64_t i = i0 + u * stride;
      ok[u] = i < total;
      q[u] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (ok[u]) {
        const int64_t r = i / cpr;
        const int64_t src = idx[r];
        if (src >= 0 && src < aug_rows) q[u] = ld_f4(aug + src * D + (int64_t)(i - r * cpr) * 4);
        tt[u] = ld_f4(t + i * 4);
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      if (!ok[u]) continue;
      const int64_t i = i0 + u * stride;
      if (q_out) st_f4(q_out + i * 4, q[u]);
      st_f4(o + i * 4, make_float4(tt[u].x + q[u].x, tt[u].y + q[u].y, tt[u].z + q[u].z, tt[u].w + q[u].w));
    }
  }
}

static inline bool al16_all(std::initializer_list<const void*> ps) {
  for (const void* p : ps)
    if ((uintptr_t)p & 15) return false;
  return true;
}

// ------------------------------------------------------------------------------------------------
// fused loss: one warp per positive pair
// ------------------------------------------------------------------------------------------------
constexpr int kMaxNeg = 16;

__device__ __forceinline__ float softplusf_(float x) { return fmaxf(x, 0.f) + log1pf(expf(-fabsf(x))); }

__global__ void __launch_bounds__(256) loss_kernel(const float* __restrict__ o_u, const float* __restrict__ o_i,
                                                   const float* __restrict__ t_u, const float* __restrict__ t_p,
                                                   const float* __restrict__ q_u, const float* __restrict__ q_p,
                                                   float cu, float ci, float* __restrict__ partial,
                                                   float* __restrict__ do_u, float* __restrict__ do_i,
                                                   float* __restrict__ dq_u, float* __restrict__ dq_p, int B, int N,
                                                   int D, float inv_M) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= B) return;
  const int b = warp;
  const float* u = o_u + (int64_t)b * D;
  const float* p = o_i + (int64_t)b * D;
  const float* n0 = o_i + ((int64_t)B + (int64_t)b * N) * D;
  float sp = 0.f;
  float sn[kMaxNeg];
#pragma unroll
  for (int n = 0; n < kMaxNeg; ++n) sn[n] = 0.f;
  for (int d = lane; d < D; d += 32) {
    float uu = u[d];
    sp = fmaf(uu, p[d], sp);
#pragma unroll
    for (int n = 0; n < kMaxNeg; ++n)
      if (n < N) sn[n] = fmaf(uu, n0[(int64_t)n * D + d], sn[n]);
  }
  sp = warp_sum(sp);
  float bce = softplusf_(-sp);
  float dsp = (sigmoidf_(sp) - 1.f) * inv_M;
#pragma unroll
  for (int n = 0; n < kMaxNeg; ++n)
    if (n < N) {
      float s = warp_sum(sn[n]);
      bce += softplusf_(s);
      sn[n] = sigmoidf_(s) * inv_M;  // reuse as ds_neg
    }
  float mu = 0.f, mi = 0.f;
  const bool mimic = (q_u != nullptr);
  const bool bwd = (do_u != nullptr);
  if (bwd || mimic) {
    for (int d = lane; d < D; d += 32) {
      float uu = u[d], pp = p[d];
      float gu = dsp * pp;
      if (bwd) {
#pragma unroll
        for (int n = 0; n < kMaxNeg; ++n)
          if (n < N) {
            float nn = n0[(int64_t)n * D + d];
            gu = fmaf(sn[n], nn, gu);
            do_i[((int64_t)B + (int64_t)b * N + n) * D + d] = sn[n] * uu;
          }
        do_u[(int64_t)b * D + d] = gu;
        do_i[(int64_t)b * D + d] = dsp * uu;
      }
      if (mimic) {
        float du = q_u[(int64_t)b * D + d] - t_p[(int64_t)b * D + d];
        float di = q_p[(int64_t)b * D + d] - t_u[(int64_t)b * D + d];
        mu = fmaf(du, du, mu);
        mi = fmaf(di, di, mi);
        if (bwd) {
          dq_u[(int64_t)b * D + d] = fmaf(cu, du, gu);
          dq_p[(int64_t)b * D + d] = fmaf(ci, di, dsp * uu);
        }
      }
    }
    mu = warp_sum(mu);
    mi = warp_sum(mi);
  }
  if (lane == 0) {
    partial[(int64_t)b * 3 + 0] = bce;
    partial[(int64_t)b * 3 + 1] = mu;
    partial[(int64_t)b * 3 + 2] = mi;
  }
}

// 16-byte form for D <= 128, D % 4 == 0 (every BASELINE config with D in {96, 128}): lane l owns columns [4l, 4l+4).
// Every row the pair needs - u, p, the N negatives, and the four mimic rows - is fetched with ONE float4 load per lane, all
// issued before the first use (up to 2 + N + 4 loads in flight per lane), kept in registers for the gradient pass (the
// scalar kernel re-read them), and every gradient row leaves as one float4 store per lane.
template <int NMAX>
__global__ void __launch_bounds__(256) loss_vec_kernel(const float* __restrict__ o_u, const float* __restrict__ o_i,
                                                       const float* __restrict__ t_u, const float* __restrict__ t_p,
                                                       const float* __restrict__ q_u, const float* __restrict__ q_p,
                                                       float cu, float ci, float* __restrict__ partial,
                                                       float* __restrict__ do_u, float* __restrict__ do_i,
                                                       float* __restrict__ dq_u, float* __restrict__ dq_p, int B, int N,
                                                       int D, float inv_M) {
  const int b = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (b >= B) return;
  const int col = lane * 4;
  const bool act = col < D;
  const bool mimic = (q_u != nullptr);
  const bool bwd = (do_u != nullptr);
  const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
  const int64_t rb = (int64_t)b * D + col;                                  // row b of a [B, D] block
  const int64_t rn = ((int64_t)B + (int64_t)b * N) * D + col;               // first negative of b in o_i / do_i
  float4 uu = z4, pp = z4, nn[NMAX], qu = z4, tp = z4, qp = z4, tu = z4;
#pragma unroll
  for (int n = 0; n < NMAX; ++n) nn[n] = z4;
  if (act) {
    uu = ld_f4(o_u + rb);
    pp = ld_f4(o_i + rb);
#pragma unroll
    for (int n = 0; n < NMAX; ++n)
      if (n < N) nn[n] = ld_f4(o_i + rn + (int64_t)n * D);
    if (mimic) {
      qu = ld_f4(q_u + rb); tp = ld_f4(t_p + rb); qp = ld_f4(q_p + rb); tu = ld_f4(t_u + rb);
    }
  }
  auto dot4 = [](const float4& a, const float4& c) { return fmaf(a.w, c.w, fmaf(a.z, c.z, fmaf(a.y, c.y, a.x * c.x))); };
  const float sp = warp_sum(dot4(uu, pp));
  float bce = softplusf_(-sp);
  const float dsp = (sigmoidf_(sp) - 1.f) * inv_M;
  float ds[NMAX];
#pragma unroll
  for (int n = 0; n < NMAX; ++n) {
    ds[n] = 0.f;
    if (n < N) {
      const float sc = warp_sum(dot4(uu, nn[n]));
      bce += softplusf_(sc);
      ds[n] = sigmoidf_(sc) * inv_M;
    }
  }
  float mu = 0.f, mi = 0.f;
  float4 gu = make_float4(dsp * pp.x, dsp * pp.y, dsp * pp.z, dsp * pp.w);
  const float4 gp = make_float4(dsp * uu.x, dsp * uu.y, dsp * uu.z, dsp * uu.w);
  if (bwd) {
#pragma unroll
    for (int n = 0; n < NMAX; ++n)
      if (n < N) {
        gu.x = fmaf(ds[n], nn[n].x, gu.x); gu.y = fmaf(ds[n], nn[n].y, gu.y);
        gu.z = fmaf(ds[n], nn[n].z, gu.z); gu.w = fmaf(ds[n], nn[n].w, gu.w);
        if (act) st_f4(do_i + rn + (int64_t)n * D, make_float4(ds[n] * uu.x, ds[n] * uu.y, ds[n] * uu.z, ds[n] * uu.w));
      }
    if (act) {
      st_f4(do_u + rb, gu);
      st_f4(do_i + rb, gp);
    }
  }
  if (mimic) {
    const float4 du = make_float4(qu.x - tp.x, qu.y - tp.y, qu.z - tp.z, qu.w - tp.w);
    const float4 di = make_float4(qp.x - tu.x, qp.y - tu.y, qp.z - tu.z, qp.w - tu.w);
    mu = warp_sum(dot4(du, du));
    mi = warp_sum(dot4(di, di));
    if (bwd && act) {
      st_f4(dq_u + rb, make_float4(fmaf(cu, du.x, gu.x), fmaf(cu, du.y, gu.y), fmaf(cu, du.z, gu.z), fmaf(cu, du.w, gu.w)));
      st_f4(dq_p + rb, make_float4(fmaf(ci, di.x, gp.x), fmaf(ci, di.y, gp.y), fmaf(ci, di.z, gp.z), fmaf(ci, di.w, gp.w)));
    }
  }
  if (lane == 0) {
    partial[(int64_t)b * 3 + 0] = bce;
    partial[(int64_t)b * 3 + 1] = mu;
    partial[(int64_t)b * 3 + 2] = mi;
  }
}

// The same loss with the augmentation add folded in (BASELINE north_star (2)): o = t + A[idx] is formed in registers from
// the base tower rows t and the augmentation rows of the step's ids, so o and q never exist in HBM and the separate
// augment launch (read t, gather A, write o and q) disappears from the step.  Indices of the pair's 2 + N rows are one
// load by the first lanes; every row of the pair (2 + N base rows, 2 + N augmentation rows) is requested before the
// first use.  Arithmetic identical to ttam_augment_fwd followed by loss_vec_kernel (o = t + q is the same fp32 add).
template <int NMAX>
__global__ void __launch_bounds__(256) loss_aug_vec_kernel(const float* __restrict__ t_u, const float* __restrict__ t_i,
                                                           const float* __restrict__ A_u, int64_t rows_u,
                                                           const float* __restrict__ A_i, int64_t rows_i,
                                                           const int64_t* __restrict__ users, const int64_t* __restrict__ items,
                                                           int mimic, float cu, float ci, float* __restrict__ partial,
                                                           float* __restrict__ do_u, float* __restrict__ do_i,
                                                           float* __restrict__ dq_u, float* __restrict__ dq_p, int B, int N,
                                                           int D, float inv_M) {
  const int b = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (b >= B) return;
  const int col = lane * 4;
  const bool act = col < D;
  const bool bwd = (do_u != nullptr);
  const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
  int64_t my = -1;   // lane 0: the user, lane 1: the positive, lanes 2 .. 1+N: the negatives
  if (lane == 0) my = users[b];
  else if (lane == 1) my = items[b];
  else if (lane < 2 + N) my = items[(int64_t)B + (int64_t)b * N + (lane - 2)];
  if (my >= (lane == 0 ? rows_u : rows_i)) my = -1;
  const int64_t rb = (int64_t)b * D + col;                                  // row b of a [B, D] block
  const int64_t rn = ((int64_t)B + (int64_t)b * N) * D + col;               // first negative of b in t_i / do_i
  const int64_t iu = __shfl_sync(0xffffffffu, my, 0), ip = __shfl_sync(0xffffffffu, my, 1);
  int64_t in_[NMAX];
#pragma unroll
  for (int n = 0; n < NMAX; ++n) in_[n] = __shfl_sync(0xffffffffu, my, 2 + n);
  float4 tu = z4, tp = z4, qu = z4, qp = z4, tn[NMAX], qn[NMAX];
#pragma unroll
  for (int n = 0; n < NMAX; ++n) tn[n] = qn[n] = z4;
  if (act) {
    tu = ld_f4(t_u + rb);
    tp = ld_f4(t_i + rb);
    if (iu >= 0) qu = ld_f4(A_u + iu * D + col);
    if (ip >= 0) qp = ld_f4(A_i + ip * D + col);
#pragma unroll
    for (int n = 0; n < NMAX; ++n)
      if (n < N) {
        tn[n] = ld_f4(t_i + rn + (int64_t)n * D);
        if (in_[n] >= 0) qn[n] = ld_f4(A_i + in_[n] * D + col);
      }
  }
  auto add4 = [](const float4& a, const float4& c) { return make_float4(a.x + c.x, a.y + c.y, a.z + c.z, a.w + c.w); };
  const float4 uu = add4(tu, qu), pp = add4(tp, qp);
  float4 nn[NMAX];
#pragma unroll
  for (int n = 0; n < NMAX; ++n) nn[n] = add4(tn[n], qn[n]);
  auto dot4 = [](const float4& a, const float4& c) { return fmaf(a.w, c.w, fmaf(a.z, c.z, fmaf(a.y, c.y, a.x * c.x))); };
  const float sp = warp_sum(dot4(uu, pp));
  float bce = softplusf_(-sp);
  const float dsp = (sigmoidf_(sp) - 1.f) * inv_M;
  float ds[NMAX];
#pragma unroll
  for (int n = 0; n < NMAX; ++n) {
    ds[n] = 0.f;
    if (n < N) {
      const float sc = warp_sum(dot4(uu, nn[n]));
      bce += softplusf_(sc);
      ds[n] = sigmoidf_(sc) * inv_M;
    }
  }
  float mu = 0.f, mi = 0.f;
  float4 gu = make_float4(dsp * pp.x, dsp * pp.y, dsp * pp.z, dsp * pp.w);
  const float4 gp = make_float4(dsp * uu.x, dsp * uu.y, dsp * uu.z, dsp * uu.w);
  if (bwd) {
#pragma unroll
    for (int n = 0; n < NMAX; ++n)
      if (n < N) {
        gu.x = fmaf(ds[n], nn[n].x, gu.x); gu.y = fmaf(ds[n], nn[n].y, gu.y);
        gu.z = fmaf(ds[n], nn[n].z, gu.z); gu.w = fmaf(ds[n], nn[n].w, gu.w);
        if (act) st_f4(do_i + rn + (int64_t)n * D, make_float4(ds[n] * uu.x, ds[n] * uu.y, ds[n] * uu.z, ds[n] * uu.w));
      }
    if (act) {
      st_f4(do_u + rb, gu);
      st_f4(do_i + rb, gp);
    }
  }
  if (mimic) {
    const float4 du = make_float4(qu.x - tp.x, qu.y - tp.y, qu.z - tp.z, qu.w - tp.w);
    const float4 di = make_float4(qp.x - tu.x, qp.y - tu.y, qp.z - tu.z, qp.w - tu.w);
    mu = warp_sum(dot4(du, du));
    mi = warp_sum(dot4(di, di));
    if (bwd && act) {
      st_f4(dq_u + rb, make_float4(fmaf(cu, du.x, gu.x), fmaf(cu, du.y, gu.y), fmaf(cu, du.z, gu.z), fmaf(cu, du.w, gu.w)));
      st_f4(dq_p + rb, make_float4(fmaf(ci, di.x, gp.x), fmaf(ci, di.y, gp.y), fmaf(ci, di.z, gp.z), fmaf(ci, di.w, gp.w)));
    }
  }
  if (lane == 0) {
    partial[(int64_t)b * 3 + 0] = bce;
    partial[(int64_t)b * 3 + 1] = mu;
    partial[(int64_t)b * 3 + 2] = mi;
  }
}

// The loss of the row-sharded step with the row exchange folded in (SURVEY 8(e) all-to-all #2 and #3 as the loads and
// stores of ONE kernel over NVLink peer mappings): a pair's base rows t and augmentation rows q are loaded straight from
// the buffers of the ranks that OWN them (slot s of an index set = owner s / cap, position s % cap, as ttam_slot_plan
// laid them out), o = t + q is formed in registers, and the gradient rows [do | dq] are stored straight into the
// owners' receive buffers at the same slots.  Replaces ttam_slot_unpack -> ttam_loss_fwd_bwd -> ttam_slot_pack (three
// passes over the rows, two of them only to move data); padding slots are never written - the owner zeroes its receive
// buffers at the start of the step.
constexpr int kSlotMaxW = 16;
struct SlotLossPtrs {
  const float* t_u[kSlotMaxW]; const float* q_u[kSlotMaxW]; const float* t_i[kSlotMaxW]; const float* q_i[kSlotMaxW];
  float* a_u[kSlotMaxW]; float* b_u[kSlotMaxW]; float* a_i[kSlotMaxW]; float* b_i[kSlotMaxW];
};

template <int NMAX>
__global__ void __launch_bounds__(256) loss_slots_vec_kernel(const SlotLossPtrs P, int W, int64_t cap_u, int64_t cap_i,
                                                             const int64_t* __restrict__ slot_of_u,
                                                             const int64_t* __restrict__ slot_of_i, float cu, float ci,
                                                             float* __restrict__ partial, int B, int N, int D, float inv_M) {
  const int b = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (b >= B) return;
  const int col = lane * 4;
  const bool act = col < D;
  const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
  // lane 0: the user's slot, lane 1: the positive's, lanes 2 .. 1+N: the negatives'; owner and element offset per lane
  int64_t slot = -1;
  if (lane == 0) slot = slot_of_u[b];
  else if (lane == 1) slot = slot_of_i[b];
  else if (lane < 2 + N) slot = slot_of_i[(int64_t)B + (int64_t)b * N + (lane - 2)];
  const int64_t cap = lane == 0 ? cap_u : cap_i;
  int own = -1;
  int64_t off = 0;
  if (slot >= 0 && slot < (int64_t)W * cap) {   // (a slot of W * cap = "did not fit": the caller routes such steps elsewhere)
    own = (int)(slot / cap);
    off = (slot - (int64_t)own * cap) * D;
  }
  const int ou = __shfl_sync(0xffffffffu, own, 0), op = __shfl_sync(0xffffffffu, own, 1);
  const int64_t eu_ = __shfl_sync(0xffffffffu, off, 0) + col, ep_ = __shfl_sync(0xffffffffu, off, 1) + col;
  int on[NMAX];
  int64_t en[NMAX];
#pragma unroll
  for (int n = 0; n < NMAX; ++n) {
    on[n] = __shfl_sync(0xffffffffu, own, 2 + n);
    en[n] = __shfl_sync(0xffffffffu, off, 2 + n) + col;
  }
  float4 tu = z4, tp = z4, qu = z4, qp = z4, tn[NMAX], qn[NMAX];
#pragma unroll
  for (int n = 0; n < NMAX; ++n) tn[n] = qn[n] = z4;
  if (act) {
    if (ou >= 0) { tu = ld_f4_stream(P.t_u[ou] + eu_); qu = ld_f4_stream(P.q_u[ou] + eu_); }
    if (op >= 0) { tp = ld_f4_stream(P.t_i[op] + ep_); qp = ld_f4_stream(P.q_i[op] + ep_); }
#pragma unroll
    for (int n = 0; n < NMAX; ++n)
      if (n < N && on[n] >= 0) {
        tn[n] = ld_f4_stream(P.t_i[on[n]] + en[n]);
        qn[n] = ld_f4_stream(P.q_i[on[n]] + en[n]);
      }
  }
  auto add4 = [](const float4& a, const float4& c) { return make_float4(a.x + c.x, a.y + c.y, a.z + c.z, a.w + c.w); };
  const float4 uu = add4(tu, qu), pp = add4(tp, qp);
  float4 nn[NMAX];
#pragma unroll
  for (int n = 0; n < NMAX; ++n) nn[n] = add4(tn[n], qn[n]);
  auto dot4 = [](const float4& a, const float4& c) { return fmaf(a.w, c.w, fmaf(a.z, c.z, fmaf(a.y, c.y, a.x * c.x))); };
  const float sp = warp_sum(dot4(uu, pp));
  float bce = softplusf_(-sp);
  const float dsp = (sigmoidf_(sp) - 1.f) * inv_M;
  float4 gu = make_float4(dsp * pp.x, dsp * pp.y, dsp * pp.z, dsp * pp.w);
  const float4 gp = make_float4(dsp * uu.x, dsp * uu.y, dsp * uu.z, dsp * uu.w);
#pragma unroll
  for (int n = 0; n < NMAX; ++n) {
    if (n < N) {
      const float sc = warp_sum(dot4(uu, nn[n]));
      bce += softplusf_(sc);
      const float dsn = sigmoidf_(sc) * inv_M;
      gu.x = fmaf(dsn, nn[n].x, gu.x); gu.y = fmaf(dsn, nn[n].y, gu.y);
      gu.z = fmaf(dsn, nn[n].z, gu.z); gu.w = fmaf(dsn, nn[n].w, gu.w);
      if (act && on[n] >= 0) {
        const float4 g = make_float4(dsn * uu.x, dsn * uu.y, dsn * uu.z, dsn * uu.w);
        st_f4(P.a_i[on[n]] + en[n], g);       // do of a negative ...
        st_f4(P.b_i[on[n]] + en[n], g);       // ... is also the gradient of its augmentation row
      }
    }
  }
  const float4 du = make_float4(qu.x - tp.x, qu.y - tp.y, qu.z - tp.z, qu.w - tp.w);
  const float4 di = make_float4(qp.x - tu.x, qp.y - tu.y, qp.z - tu.z, qp.w - tu.w);
  const float mu = warp_sum(dot4(du, du));
  const float mi = warp_sum(dot4(di, di));
  if (act) {
    if (ou >= 0) {
      st_f4(P.a_u[ou] + eu_, gu);
      st_f4(P.b_u[ou] + eu_, make_float4(fmaf(cu, du.x, gu.x), fmaf(cu, du.y, gu.y), fmaf(cu, du.z, gu.z), fmaf(cu, du.w, gu.w)));
    }
    if (op >= 0) {
      st_f4(P.a_i[op] + ep_, gp);
      st_f4(P.b_i[op] + ep_, make_float4(fmaf(ci, di.x, gp.x), fmaf(ci, di.y, gp.y), fmaf(ci, di.z, gp.z), fmaf(ci, di.w, gp.w)));
    }
  }
  if (lane == 0) {
    partial[(int64_t)b * 3 + 0] = bce;
    partial[(int64_t)b * 3 + 1] = mu;
    partial[(int64_t)b * 3 + 2] = mi;
  }
}

// deterministic final reduction: one block, fixed-order strided partial sums + tree
__global__ void __launch_bounds__(1024) loss_reduce_kernel(const float* __restrict__ partial, int B, float inv_M,
                                                           float inv_BD, float lambda_u, float lambda_i, int mimic,
                                                           float* __restrict__ loss_out) {
  __shared__ double sh[3][1024];
  double a0 = 0, a1 = 0, a2 = 0;
  for (int b = threadIdx.x; b < B; b += blockDim.x) {
    a0 += partial[(int64_t)b * 3 + 0];
    a1 += partial[(int64_t)b * 3 + 1];
    a2 += partial[(int64_t)b * 3 + 2];
  }
  sh[0][threadIdx.x] = a0;
  sh[1][threadIdx.x] = a1;
  sh[2][threadIdx.x] = a2;
  __syncthreads();
  for (int s = blockDim.x >> 1; s > 0; s >>= 1) {
    if (threadIdx.x < s) {
      sh[0][threadIdx.x] += sh[0][threadIdx.x + s];
      sh[1][threadIdx.x] += sh[1][threadIdx.x + s];
      sh[2][threadIdx.x] += sh[2][threadIdx.x + s];
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    float bce = (float)(sh[0][0] * (double)inv_M);
    float mu = (float)(sh[1][0] * (double)inv_BD);
    float mi = (float)(sh[2][0] * (double)inv_BD);
    float total = bce;
    if (mimic && lambda_u > 0.f) total += lambda_u * mu;
    if (mimic && lambda_i > 0.f) total += lambda_i * mi;
    loss_out[0] = total;
    loss_out[1] = bce;
    loss_out[2] = mu;
    loss_out[3] = mi;
  }
}

// ------------------------------------------------------------------------------------------------
// step state, elementwise activation (+dropout) for the non-ReLU feature encoders
// ------------------------------------------------------------------------------------------------
__global__ void advance_step_kernel(ttam_step_state* st, uint64_t rng_stride) {
  st->step += 1;
  st->rng_offset += rng_stride;
}

__device__ __forceinline__ float act_fwd_(int act, float x) {
  switch (act) {
    case TTAM_ACT_RELU: return fmaxf(x, 0.f);
    case TTAM_ACT_TANH: return tanhf(x);
    case TTAM_ACT_GELU: return 0.5f * x * (1.f + erff(x * 0.70710678118654752440f));
    case TTAM_ACT_SELU: {
      const float alpha = 1.6732632423543772848170429916717f, scale = 1.0507009873554804934193349852946f;
      return scale * (x > 0.f ? x : alpha * (expf(x) - 1.f));
    }
    default: return x;
  }
}
__device__ __forceinline__ float act_bwd_(int act, float x) {
  switch (act) {
    case TTAM_ACT_RELU: return x > 0.f ? 1.f : 0.f;
    case TTAM_ACT_TANH: { float t = tanhf(x); return 1.f - t * t; }
    case TTAM_ACT_GELU: {
      const float cdf = 0.5f * (1.f + erff(x * 0.70710678118654752440f));
      const float pdf = expf(-0.5f * x * x) * 0.39894228040143267794f;
      return cdf + x * pdf;
    }
    case TTAM_ACT_SELU: {
      const float alpha = 1.6732632423543772848170429916717f, scale = 1.0507009873554804934193349852946f;
      return x > 0.f ? scale : scale * alpha * expf(x);
    }
    default: return 1.f;
  }
}

template <bool BWD>
__global__ void __launch_bounds__(256) act_kernel(const float* __restrict__ dy, const float* __restrict__ pre,
                                                  float* __restrict__ out, int64_t n, int act, float p, uint64_t seed,
                                                  uint64_t offset, const ttam_step_state* __restrict__ st) {
  const float keep_scale = p > 0.f ? 1.f / (1.f - p) : 1.f;
  const uint64_t base = offset + ((p > 0.f && st) ? st->rng_offset : 0ull);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float x = pre[i];
    float v = BWD ? dy[i] * act_bwd_(act, x) : act_fwd_(act, x);
    if (p > 0.f) v = dropout_keep(seed, base + (uint64_t)i, p) ? v * keep_scale : 0.f;
    out[i] = v;
  }
}

static inline int grid_for(int64_t total, int threads, int per_thread = 4) {
  int64_t blocks = ceil_div(total, (int64_t)threads * per_thread);
  int64_t cap = (int64_t)num_sms() * 16;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (int)blocks;
}

}  // namespace ttam

using namespace ttam;

extern "C" int ttam_gather_rows_f32(const float* table, int64_t ld_table, int64_t num_rows, const int64_t* idx,
                                    float* out, int64_t ld_out, int64_t R, int64_t ncols, void* stream) {
  TTAM_CHECK_ARG(R >= 0 && ncols > 0 && ld_table >= ncols && ld_out >= ncols, "gather_rows: bad shape");
  if (R == 0) return TTAM_OK;  // an empty batch has no rows to address (its pointers may be null)
  TTAM_CHECK_ARG(table && idx && out, "gather_rows: null pointer");
  cudaStream_t s = (cudaStream_t)stream;
  bool vec = (ncols % 4 == 0) && (ld_table % 4 == 0) && (ld_out % 4 == 0) && (((uintptr_t)table & 15) == 0) &&
             (((uintptr_t)out & 15) == 0);
  if (vec)
    gather_rows_kernel<true><<<(unsigned)ceil_div(R, (int64_t)kGatherRows * 8), 256, 0, s>>>(table, ld_table, num_rows, idx, out,
                                                                                             ld_out, R, ncols);
  else
    gather_rows_kernel<false><<<grid_for(R * ncols, 256), 256, 0, s>>>(table, ld_table, num_rows, idx, out, ld_out,
                                                                      R, ncols);
  TTAM_LAUNCH_CHECK();
  return TTAM_OK;
}

__global__ void __launch_bounds__(256) round_tf32_kernel(float* __restrict__ x, int64_t ld, int64_t R, int64_t ncols) {
  const int64_t total = R * ncols;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / ncols, c = i - r * ncols;
    uint32_t u;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x[r * ld + c]));
    x[r * ld + c] = __uint_as_float(u);
  }
}

extern "C" int ttam_round_tf32(float* x, int64_t ld, int64_t R, int64_t ncols, void* stream) {
  TTAM_CHECK_ARG(R >= 0 && ncols >= 0 && ld >= ncols, "round_tf32: bad shape");
  if (R == 0 || ncols == 0) return TTAM_OK;
  TTAM_CHECK_ARG(x, "round_tf32: null pointer");
  const int blocks = (int)std::min<int64_t>(ceil_div(R * ncols, 256), (int64_t)num_sms() * 16);
  round_tf32_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(x, ld, R, ncols);
  TTAM_LAUNCH_CHECK();
  return TTAM_OK;
}

extern "C" int ttam_cast_f32_to_bf16(const float* src, int64_t ld_src, uint16_t* dst, int64_t ld_dst, int64_t R,
                                     int64_t ncols, void* stream) {
  TTAM_CHECK_ARG(src && dst && R >= 0 && ncols > 0, "cast: bad argument");
  if (R == 0) return TTAM_OK;
  cast_bf16_kernel<<<grid_for(R * ncols, 256), 256, 0, (cudaStream_t)stream>>>(src, ld_src, dst, ld_dst, R, ncols);
  TTAM_LAUNCH_CHECK();
  return TTAM_OK;
}

extern "C" int ttam_gate_fwd(const float* z, const float* pre2, const float* aug_table, int64_t aug_rows,
                             const int64_t* idx, float* g, float* t, float* o, float* q_out, int64_t R, int64_t D,
                             void* stream) {
  TTAM_CHECK_ARG(z && pre2 && g, "gate_fwd: null pointer");
  TTAM_CHECK_ARG(!aug_table || idx, "gate_fwd: augmentation needs indices");
  if (R == 0) return TTAM_OK;
  if (D % 4 == 0 && D < (1 << 20) && al16_all({z, pre2, aug_table, g, t, o, q_out}))
    gate_fwd_vec_kernel<<<grid_for(R * (D / 4), 256, 2), 256, 0, (cudaStream_t)stream>>>(z, pre2, aug_table, aug_rows, idx, g,
                                                                                        t, o, q_out, R, (int)D);
  else
    gate_fwd_kernel<<<grid_for(R * D, 256, 2), 256, 0, (cudaStream_t)stream>>>(z, pre2, aug_table, aug_rows, idx, g, t,
                                                                               o, q_out, R, D);
  TTAM_LAUNCH_CHECK();
  return TTAM_OK;
}

extern "C" int ttam_gate_bwd(const float* dt, const float* z, const float* g, float* dpre2, float* dz, int64_t R,
                             int64_t D, void* stream) {
  TTAM_CHECK_ARG(dt && z && g && dpre2 && dz, "gate_bwd: null pointer");
  if (R == 0) return TTAM_OK;
  if (D % 4 == 0 && D < (1 << 20) && al16_all({dt, z, g, dpre2, dz}))
    gate_bwd_vec_kernel<<<grid_for(R * (D / 4), 256, 2), 256, 0, (cudaStream_t)stream>>>(dt, z, g, dpre2, dz, R, (int)D);
  else
    gate_bwd_kernel<<<grid_for(R * D, 256, 2), 256, 0, (cudaStream_t)stream>>>(dt, z, g, dpre2, dz, R, D);
  TTAM_LAUNCH_CHECK();
  return TTAM_OK;
}

extern "C" int ttam_augment_fwd(const float* t, const float* aug_table, int64_t aug_rows, const int64_t* idx,
                                float* o, float* q_out, int64_t R, int64_t D, void* stream) {
  TTAM_CHECK_ARG(t && aug_table && idx && o, "augment_fwd: null pointer");
  if (R == 0) return TTAM_OK;
  if (D % 4 == 0 && D < (1 << 20) && al16_all({t, aug_table, o, q_out}))
    augment_fwd_vec_kernel<<<grid_for(R * (D / 4), 256, 4), 256, 0, (cudaStream_t)stream>>>(t, aug_table, aug_rows, idx, o,
                                                                                           q_out, R, (int)D);
  else
    augment_fwd_kernel<<<grid_for(R * D, 256, 2), 256, 0, (cudaStream_t)stream>>>(t, aug_table, aug_rows, idx, o, q_out,
                                                                                  R, D);
  TTAM_LAUNCH_CHECK();
  return TTAM_OK;
}

extern "C" int64_t ttam_loss_workspace_bytes(int64_t B) { return B * 3 * (int64_t)sizeof(float); }

extern "C" int ttam_loss_fwd_bwd(const float* o_u, const float* o_i, const float* t_u, const float* t_p,
                                 const float* q_u, const float* q_p, float lambda_u, float lambda_i, float* loss_out,
                                 float* do_u, float* do_i, float* dq_u, float* dq_p, int64_t B, int64_t N, int64_t D,
                                 float batch_fraction, void* workspace, int64_t workspace_bytes, void* stream) {
  TTAM_CHECK_ARG(o_u && o_i && loss_out && workspace, "loss: null pointer");
  TTAM_CHECK_ARG(batch_fraction > 0.f && batch_fraction <= 1.f, "loss: batch_fraction must be in (0, 1]");
  TTAM_CHECK_ARG(B > 0 && N > 0 && N <= kMaxNeg && D > 0, "loss: need B>0, 0<N<=%d, D>0", kMaxNeg);
  TTAM_CHECK_ARG((q_u == nullptr) == (q_p == nullptr) && (q_u == nullptr) == (t_u == nullptr) &&
                     (q_u == nullptr) == (t_p == nullptr),
                 "loss: mimic inputs must be all set or all null");
  TTAM_CHECK_ARG((do_u == nullptr) == (do_i == nullptr), "loss: do_u/do_i must both be set or both null");
  TTAM_CHECK_ARG(!(do_u && q_u) || (dq_u && dq_p), "loss: dq_u/dq_p required with mimic backward");
  if (workspace_bytes < ttam_loss_workspace_bytes(B)) {
    set_error("loss: workspace too small");
    return TTAM_EWORKSPACE;
  }
  cudaStream_t s = (cudaStream_t)stream;
  const double M = (double)B * (double)(1 + N);
  // data-parallel ranks each hold `batch_fraction` of the global batch: the means run over the global batch
  const float inv_M = (float)((double)batch_fraction / M);
  const float inv_BD = (float)((double)batch_fraction / ((double)B * (double)D));
  // dL/dq extra term: lambda * 2 (q - t) / (B*D)   (only when lambda > 0: training.py:800-803)
  const float cu = lambda_u > 0.f ? (float)(2.0 * (double)lambda_u * (double)batch_fraction / ((double)B * (double)D)) : 0.f;
  const float ci = lambda_i > 0.f ? (float)(2.0 * (double)lambda_i * (double)batch_fraction / ((double)B * (double)D)) : 0.f;
  const int threads = 256;
  const int blocks = (int)ceil_div(B * 32, threads);
  auto al16 = [](const void* p) { return ((uintptr_t)p & 15) == 0; };
  const bool vec = D % 4 == 0 && D <= 128 && N <= 8 && al16(o_u) && al16(o_i) && al16(t_u) && al16(t_p) && al16(q_u) &&
                   al16(q_p) && al16(do_u) && al16(do_i) && al16(dq_u) && al16(dq_p);
  if (vec && N <= 5)
    loss_vec_kernel<5><<<blocks, threads, 0, s>>>(o_u, o_i, t_u, t_p, q_u, q_p, cu, ci, (float*)workspace, do_u, do_i, dq_u,
                                                  dq_p, (int)B, (int)N, (int)D, inv_M);
  else if (vec)
    loss_vec_kernel<8><<<blocks, threads, 0, s>>>(o_u, o_i, t_u, t_p, q_u, q_p, cu, ci, (float*)workspace, do_u, do_i, dq_u,
                                                  dq_p, (int)B, (int)N, (int)D, inv_M);
  else
    loss_kernel<<<blocks, threads, 0, s>>>(o_u, o_i, t_u, t_p, q_u, q_p, cu, ci, (float*)workspace, do_u, do_i, dq_u,
                                           dq_p, (int)B, (int)N, (int)D, inv_M);
  TTAM_LAUNCH_CHECK();
  loss_reduce_kernel<<<1, 1024, 0, s>>>((const float*)workspace, (int)B, inv_M, inv_BD, lambda_u, lambda_i,
                                        q_u != nullptr, loss_out);
  TTAM_LAUNCH_CHECK();
  return TTAM_OK;
}

extern "C" int ttam_loss_aug_supported(int64_t N, int64_t D) { return (D % 4 == 0 && D <= 128 && N >= 1 && N <= 8) ? 1 : 0; }

extern "C" int ttam_loss_aug_fwd_bwd(const float* t_u, const float* t_i, const float* aug_u, int64_t aug_u_rows, const float* aug_i,
                                     int64_t aug_i_rows, const int64_t* users, const int64_t* items, int mimic, float lambda_u,
                                     float lambda_i, float* loss_out, float* do_u, float* do_i, float* dq_u, float* dq_p, int64_t B,
                                     int64_t N, int64_t D, float batch_fraction, void* workspace, int64_t workspace_bytes,
                                     void* stream) {
  TTAM_CHECK_ARG(t_u && t_i && aug_u && aug_i && users && items && loss_out && workspace, "loss_aug: null pointer");
  TTAM_CHECK_ARG(batch_fraction > 0.f && batch_fraction <= 1.f, "loss_aug: batch_fraction must be in (0, 1]");
  TTAM_CHECK_ARG(B > 0 && ttam_loss_aug_supported(N, D), "loss_aug: need B > 0, 1 <= N <= 8, D %% 4 == 0, D <= 128 (use ttam_augment_fwd + ttam_loss_fwd_bwd)");
  TTAM_CHECK_ARG((do_u == nullptr) == (do_i == nullptr), "loss_aug: do_u/do_i must both be set or both null");
  TTAM_CHECK_ARG(!(do_u && mimic) || (dq_u && dq_p), "loss_aug: dq_u/dq_p required with mimic backward");
  auto al16 = [](const void* p) { return ((uintptr_t)p & 15) == 0; };
  TTAM_CHECK_ARG(al16(t_u) && al16(t_i) && al16(aug_u) && al16(aug_i) && al16(do_u) && al16(do_i) && al16(dq_u) && al16(dq_p),
                 "loss_aug: 16-byte aligned rows required");
  if (workspace_bytes < ttam_loss_workspace_bytes(B)) {
    set_error("loss_aug: workspace too small");
    return TTAM_EWORKSPACE;
  }
  cudaStream_t s = (cudaStream_t)stream;
  const double M = (double)B * (double)(1 + N);
  const float inv_M = (float)((double)batch_fraction / M);
  const float inv_BD = (float)((double)batch_fraction / ((double)B * (double)D));
  const float cu = lambda_u > 0.f ? (float)(2.0 * (double)lambda_u * (double)batch_fraction / ((double)B * (double)D)) : 0.f;
  const float ci = lambda_i > 0.f ? (float)(2.0 * (double)lambda_i * (double)batch_fraction / ((double)B * (double)D)) : 0.f;
  const int threads = 256;
  const int blocks = (int)ceil_div(B * 32, threads);
  if (N <= 5)
    loss_aug_vec_kernel<5><<<blocks, threads, 0, s>>>(t_u, t_i, aug_u, aug_u_rows, aug_i, aug_i_rows, users, items, mimic, cu, ci,
                                                      (float*)workspace, do_u, do_i, dq_u, dq_p, (int)B, (int)N, (int)D, inv_M);
  else
    loss_aug_vec_kernel<8><<<blocks, threads, 0, s>>>(t_u, t_i, aug_u, aug_u_rows, aug_i, aug_i_rows, users, items, mimic, cu, ci,
                                                      (float*)workspace, do_u, do_i, dq_u, dq_p, (int)B, (int)N, (int)D, inv_M);
  TTAM_LAUNCH_CHECK();
  loss_reduce_kernel<<<1, 1024, 0, s>>>((const float*)workspace, (int)B, inv_M, inv_BD, lambda_u, lambda_i, mimic, loss_out);
  TTAM_LAUNCH_CHECK();
  return TTAM_OK;
}

extern "C" int ttam_loss_slots_fwd_bwd(const float* const* t_u, const float* const* q_u, const float* const* t_i, const float* const* q_i,
                                       float* const* a_u, float* const* b_u, float* const* a_i, float* const* b_i, int64_t world,
                                       int64_t cap_u, int64_t cap_i, const int64_t* slot_of_u, const int64_t* slot_of_i, float lambda_u,
                                       float lambda_i, float* loss_out, int64_t B, int64_t N, int64_t D, float batch_fraction,
                                       void* workspace, int64_t workspace_bytes, void* stream) {
  TTAM_CHECK_ARG(t_u && q_u && t_i && q_i && a_u && b_u && a_i && b_i && slot_of_u && slot_of_i && loss_out && workspace,
                 "loss_slots: null pointer");
  TTAM_CHECK_ARG(world >= 1 && world <= kSlotMaxW && cap_u >= 1 && cap_i >= 1, "loss_slots: world=%lld not in [1,%d] or empty slots",
                 (long long)world, kSlotMaxW);
  TTAM_CHECK_ARG(batch_fraction > 0.f && batch_fraction <= 1.f, "loss_slots: batch_fraction must be in (0, 1]");
  TTAM_CHECK_ARG(B > 0 && ttam_loss_aug_supported(N, D), "loss_slots: need B > 0, 1 <= N <= 8, D %% 4 == 0, D <= 128");
  SlotLossPtrs P{};
  for (int w = 0; w < (int)world; ++w) {
    P.t_u[w] = t_u[w]; P.q_u[w] = q_u[w]; P.t_i[w] = t_i[w]; P.q_i[w] = q_i[w];
    P.a_u[w] = a_u[w]; P.b_u[w] = b_u[w]; P.a_i[w] = a_i[w]; P.b_i[w] = b_i[w];
    const uintptr_t all = (uintptr_t)t_u[w] | (uintptr_t)q_u[w] | (uintptr_t)t_i[w] | (uintptr_t)q_i[w] | (uintptr_t)a_u[w] |
                          (uintptr_t)b_u[w] | (uintptr_t)a_i[w] | (uintptr_t)b_i[w];
    TTAM_CHECK_ARG(t_u[w] && q_u[w] && t_i[w] && q_i[w] && a_u[w] && b_u[w] && a_i[w] && b_i[w] && (all & 15) == 0,
                   "loss_slots: buffer of rank %d null or not 16-byte aligned", w);
  }
  if (workspace_bytes < ttam_loss_workspace_bytes(B)) {
    set_error("loss_slots: workspace too small");
    return TTAM_EWORKSPACE;
  }
  cudaStream_t s = (cudaStream_t)stream;
  const double M = (double)B * (double)(1 + N);
  const float inv_M = (float)((double)batch_fraction / M);
  const float inv_BD = (float)((double)batch_fraction / ((double)B * (double)D));
  const float cu = lambda_u > 0.f ? (float)(2.0 * (double)lambda_u * (double)batch_fraction / ((double)B * (double)D)) : 0.f;
  const float ci = lambda_i > 0.f ? (float)(2.0 * (double)lambda_i * (double)batch_fraction / ((double)B * (double)D)) : 0.f;
  const int threads = 256;
  const int blocks = (int)ceil_div(B * 32, threads);
  if (N <= 5)
    loss_slots_vec_kernel<5><<<blocks, threads, 0, s>>>(P, (int)world, cap_u, cap_i, slot_of_u, slot_of_i, cu, ci, (float*)workspace,
                                                        (int)B, (int)N, (int)D, inv_M);
  else
    loss_slots_vec_kernel<8><<<blocks, threads, 0, s>>>(P, (int)world, cap_u, cap_i, slot_of_u, slot_of_i, cu, ci, (float*)workspace,
                                                        (int)B, (int)N, (int)D, inv_M);
  TTAM_LAUNCH_CHECK();
  loss_reduce_kernel<<<1, 1024, 0, s>>>((const float*)workspace, (int)B, inv_M, inv_BD, lambda_u, lambda_i, 1, loss_out);
  TTAM_LAUNCH_CHECK();
  return TTAM_OK;
}

extern "C" int ttam_advance_step(ttam_step_state* state_dev, uint64_t rng_stride, void* stream) {
  TTAM_CHECK_ARG(state_dev, "advance_step: null state");
  advance_step_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(state_dev, rng_stride);
  TTAM_LAUNCH_CHECK();
  return TTAM_OK;
}

// NOTE: the dropout counter of element (m, n) is offset + m*row_len + n, the same indexing the fused
// epilogue of ttam_linear_fwd uses, so `n` must cover whole contiguous rows of length row_len.
extern "C" int ttam_act_fwd(const float* pre, float* y, int64_t n, int64_t row_len, int act, float dropout_p,
                            uint64_t seed, uint64_t offset, const ttam_step_state* state_dev, void* stream) {
  TTAM_CHECK_ARG(pre && y && n >= 0 && row_len > 0, "act_fwd: bad argument");
  TTAM_CHECK_ARG(dropout_p >= 0.f && dropout_p < 1.f, "act_fwd: dropout must be in [0,1)");
  if (n == 0) return TTAM_OK;
  act_kernel<false><<<grid_for(n, 256, 2), 256, 0, (cudaStream_t)stream>>>(nullptr, pre, y, n, act, dropout_p, seed,
                                                                          offset, state_dev);
  TTAM_LAUNCH_CHECK();
  return TTAM_OK;
}

extern "C" int ttam_act_bwd(const float* dy, const float* pre, float* dpre, int64_t n, int64_t row_len, int act,
                            float dropout_p, uint64_t seed, uint64_t offset, const ttam_step_state* state_dev,
                            void* stream) {
  TTAM_CHECK_ARG(dy && pre && dpre && n >= 0 && row_len > 0, "act_bwd: bad argument");
  if (n == 0) return TTAM_OK;
  act_kernel<true><<<grid_for(n, 256, 2), 256, 0, (cudaStream_t)stream>>>(dy, pre, dpre, n, act, dropout_p, seed, offset,
                                                                         state_dev);
  TTAM_LAUNCH_CHECK();
  return TTAM_OK;
}
