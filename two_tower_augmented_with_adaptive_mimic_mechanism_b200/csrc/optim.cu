// Sparse backward (sort + segment-reduce) fused with row-wise optimisers; no dense table gradient is
// ever materialised.
//   ttam_sort_rows        : coalesce() index work      (torch/optim/_functional.py:44)
//   ttam_sparse_adam_rows : torch.optim.SparseAdam     (torch/optim/_functional.py:24-84)
//   ttam_lazy_rows/flush  : dense AdamW/Adam/SGD semantics (torch/optim/adam.py:347-547, sgd.py) applied
//                           lazily: untouched rows replay their zero-gradient steps when next touched
//   ttam_dense_step       : the small MLP / gate tensors, one launch
// All HBM-bound: one pass over p, m, v of the touched rows, 16-byte accesses.
#include "common.cuh"
#include <cub/cub.cuh>

namespace ttam {

__global__ void iota_kernel(int32_t* v, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    v[i] = (int32_t)i;
}

static int bits_for(int64_t num_rows) {
  int b = 1;
  while (b < 63 && (1ll << b) < num_rows) ++b;
  return b;
}

static size_t sort_temp_bytes(int64_t R) {
  size_t bytes = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, bytes, (const int64_t*)nullptr, (int64_t*)nullptr, (const int32_t*)nullptr,
                                  (int32_t*)nullptr, (int)R, 0, 64);
  return bytes;
}
static size_t unique_temp_bytes(int64_t R) {
  size_t bytes = 0;
  cub::DeviceSelect::Unique(nullptr, bytes, (const int64_t*)nullptr, (int64_t*)nullptr, (int64_t*)nullptr, (int)R);
  return bytes;
}

struct GradSrc {
  const float* a;
  int64_t ld_a;
  int64_t n_a;
  const float* b;
  int64_t ld_b;
  __device__ __forceinline__ const float* row(int32_t pos) const {
    return pos < n_a ? a + (int64_t)pos * ld_a : b + ((int64_t)pos - n_a) * ld_b;
  }
};

// Sum the gradient rows of one segment [i, end) for a 128-float (VEC) / 32-float (scalar) column group.
template <bool VEC>
__device__ __forceinline__ void segment_sum(const GradSrc& gs, const int64_t* __restrict__ sorted,
                                            const int32_t* __restrict__ perm, int64_t i, int64_t R, int64_t row,
                                            int col, bool active, float (&acc)[4]) {
  acc[0] = acc[1] = acc[2] = acc[3] = 0.f;
  int64_t j = i;
  while (j < R) {
    // up to 4 rows in flight
    int32_t pos[4];
    int cnt = 0;
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      if (j + u < R && sorted[j + u] == row) {
        pos[u] = perm[j + u];
        cnt = u + 1;
      } else {
        break;
      }
    }
    if (cnt == 0) break;
    float v[4][4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      if (u < cnt && active) {
        const float* src = gs.row(pos[u]) + col;
        if (VEC) {
          float4 f = ld_f4(src);
          v[u][0] = f.x; v[u][1] = f.y; v[u][2] = f.z; v[u][3] = f.w;
        } else {
          v[u][0] = src[0]; v[u][1] = v[u][2] = v[u][3] = 0.f;
        }
      } else {
        v[u][0] = v[u][1] = v[u][2] = v[u][3] = 0.f;
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u)
      if (u < cnt) {
#pragma unroll
        for (int e = 0; e < 4; ++e) acc[e] = __fadd_rn(acc[e], v[u][e]);
      }
    j += cnt;
    if (cnt < 4) break;
  }
}

struct AdamScalars {
  float lr, wd, b1, b2, eps, momentum;
  float one_minus_b1, one_minus_b2;
  float decay;      // 1 - lr*wd   (AdamW)
  float step_size;  // SparseAdam: lr*sqrt(bc2)/bc1 for this step
  int fast_replay;  // 1: zero-gradient replay from the closed form of v (default, see replay()); 0: dense_elem step by step
};

__device__ __forceinline__ void sparse_adam_elem(float g, float& p, float& m, float& v, const AdamScalars& s) {
  // torch/optim/_functional.py:63-84
  const float dm = __fmul_rn(__fsub_rn(g, m), s.one_minus_b1);
  const float dv = __fmul_rn(__fsub_rn(__fmul_rn(g, g), v), s.one_minus_b2);
  const float numer = __fadd_rn(dm, m);
  const float denom = __fadd_rn(sqrtf(__fadd_rn(dv, v)), s.eps);
  m = __fadd_rn(m, dm);
  v = __fadd_rn(v, dv);
  p = __fadd_rn(p, __fmul_rn(-s.step_size, __fdiv_rn(numer, denom)));
}

// one dense-optimiser step on one element; g already includes nothing but the raw gradient
template <int KIND>
__device__ __forceinline__ void dense_elem(float g, float& p, float& m, float& v, const AdamScalars& s,
                                           float step_size, float bc2_sqrt) {
  if (KIND == TTAM_OPT_SGD) {
    if (s.wd != 0.f) g = fmaf(s.wd, p, g);
    if (s.momentum != 0.f) {
      m = fmaf(s.momentum, m, g);  // buffer starts at 0 => first step gives m = g (torch clones g)
      g = m;
    }
    p = fmaf(-s.lr, g, p);
    return;
  }
  if (KIND == TTAM_OPT_ADAMW) {
    if (s.wd != 0.f) p = __fmul_rn(p, s.decay);
  } else {
    if (s.wd != 0.f) g = fmaf(s.wd, p, g);
  }
  m = fmaf(s.one_minus_b1, __fsub_rn(g, m), m);                        // lerp_
  v = fmaf(__fmul_rn(s.one_minus_b2, g), g, __fmul_rn(v, s.b2));       // mul_(b2).addcmul_(g, g, 1-b2)
  const float denom = __fadd_rn(__fdiv_rn(sqrtf(v), bc2_sqrt), s.eps);
  p = fmaf(-step_size, __fdiv_rn(m, denom), p);                        // addcdiv_
}

// ---- long segments ------------------------------------------------------------------------------------------------
// A popular row (Zipf-distributed positives) can own hundreds of gradient rows of a batch; summing them with one warp
// would serialise the whole step behind that warp.  Segments longer than kLongSeg are listed by
// find_long_segments_kernel and processed by one whole block each: every warp sums a strided share of the rows with
// eight loads in flight, the sixteen partial sums are combined in warp order (a fixed order: results are deterministic),
// and warp 0 applies the optimiser.  The row kernels skip those segments when they are given the list.
constexpr int kLongSeg = 16;
constexpr int kLongWarps = 16;   // 512 threads per long segment
constexpr int kLongUnroll = 8;   // gradient rows in flight per warp

__global__ void __launch_bounds__(256) find_long_segments_kernel(const int64_t* __restrict__ sorted, int64_t R,
                                                                  int32_t* __restrict__ list) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i + kLongSeg >= R) return;
  const int64_t row = sorted[i];
  if (i > 0 && sorted[i - 1] == row) return;
  if (sorted[i + kLongSeg] == row) list[1 + atomicAdd(list, 1)] = (int32_t)i;
}

__device__ __forceinline__ bool is_long_segment(const int64_t* __restrict__ sorted, int64_t i, int64_t R, int64_t row) {
  return i + kLongSeg < R && sorted[i + kLongSeg] == row;
}

// first position after the segment that starts at i (32-ary search by one warp; every lane returns the result)
__device__ __forceinline__ int64_t segment_end(const int64_t* __restrict__ sorted, int64_t i, int64_t R, int64_t row, int lane) {
  int64_t lo = i, hi = R;  // sorted[lo] == row, (hi == R or sorted[hi] != row)
  while (hi - lo > 1) {
    const int64_t span = hi - lo;
    const int64_t stepw = (span + 31) / 32;
    const int64_t probe = lo + (int64_t)(lane + 1) * stepw;
    const bool same = probe < hi && sorted[probe] == row;
    const unsigned bal = __ballot_sync(0xffffffffu, same);
    const int k = __popc(bal);  // probes 1..k hit the row (the predicate is monotone)
    const int64_t new_lo = lo + (int64_t)k * stepw;
    const int64_t new_hi = lo + (int64_t)(k + 1) * stepw;
    lo = new_lo;
    hi = new_hi < hi ? new_hi : hi;
  }
  return hi;
}

// Block-cooperative sum of gradient rows [i, end) for one column group; the result is valid in warp 0.
template <bool VEC>
__device__ __forceinline__ void segment_sum_block(const GradSrc& gs, const int32_t* __restrict__ perm, int64_t i,
                                                  int64_t end, int col, bool active, float (*part)[128], float (&acc)[4]) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float a[4] = {0.f, 0.f, 0.f, 0.f};
  for (int64_t j0 = i + warp; j0 < end; j0 += kLongUnroll * kLongWarps) {
    float v[kLongUnroll][4];
#pragma unroll
    for (int u = 0; u < kLongUnroll; ++u) {
      const int64_t j = j0 + u * kLongWarps;
      v[u][0] = v[u][1] = v[u][2] = v[u][3] = 0.f;
      if (j < end && active) {
        const float* src = gs.row(perm[j]) + col;
        if (VEC) {
          const float4 f = ld_f4(src);
          v[u][0] = f.x; v[u][1] = f.y; v[u][2] = f.z; v[u][3] = f.w;
        } else {
          v[u][0] = src[0];
        }
      }
    }
#pragma unroll
    for (int u = 0; u < kLongUnroll; ++u)
#pragma unroll
      for (int e = 0; e < 4; ++e) a[e] = __fadd_rn(a[e], v[u][e]);
  }
#pragma unroll
  for (int e = 0; e < 4; ++e) part[warp][lane * 4 + e] = a[e];
  __syncthreads();
  if (warp == 0) {
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      float t = 0.f;
      for (int w = 0; w < kLongWarps; ++w) t = __fadd_rn(t, part[w][lane * 4 + e]);
      acc[e] = t;
    }
  }
  __syncthreads();
}

template <bool VEC>
__device__ __forceinline__ void sparse_adam_update(float* __restrict__ P, float* __restrict__ Mo, float* __restrict__ Vo,
                                                   int64_t off, const float (&g)[4], const AdamScalars& s) {
  if (VEC) {
    float4 p4 = ld_f4(P + off), m4 = ld_f4(Mo + off), v4 = ld_f4(Vo + off);
    float pp[4] = {p4.x, p4.y, p4.z, p4.w}, mm[4] = {m4.x, m4.y, m4.z, m4.w}, vv[4] = {v4.x, v4.y, v4.z, v4.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) sparse_adam_elem(g[e], pp[e], mm[e], vv[e], s);
    st_f4(P + off, make_float4(pp[0], pp[1], pp[2], pp[3]));
    st_f4(Mo + off, make_float4(mm[0], mm[1], mm[2], mm[3]));
    st_f4(Vo + off, make_float4(vv[0], vv[1], vv[2], vv[3]));
  } else {
    float pp = P[off], mm = Mo[off], vv = Vo[off];
    sparse_adam_elem(g[0], pp, mm, vv, s);
    P[off] = pp; Mo[off] = mm; Vo[off] = vv;
  }
}

template <bool VEC>
__global__ void __launch_bounds__(256, 6) sparse_adam_rows_kernel(float* __restrict__ P, float* __restrict__ Mo,
                                                               float* __restrict__ Vo, int D,
                                                               const int64_t* __restrict__ sorted,
                                                               const int32_t* __restrict__ perm, int64_t R, GradSrc gs,
                                                               AdamScalars s, const float* __restrict__ scalars,
                                                               const ttam_step_state* __restrict__ st, bool skip_long) {
  const int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (i >= R) return;
  const int64_t row = sorted[i];
  if (i > 0 && sorted[i - 1] == row) return;  // not a segment head
  if (skip_long && is_long_segment(sorted, i, R, row)) return;
  if (st) s.step_size = scalars[4 * st->step + 2];
  constexpr int W = VEC ? 128 : 32;
  for (int c0 = 0; c0 < D; c0 += W) {
    const int col = c0 + (VEC ? lane * 4 : lane);
    const bool active = col < D;
    // the row's p, m, v do not depend on the gradient rows: fetch them first so that both latency chains overlap
    const int64_t off = row * (int64_t)D + col;
    float4 p4 = make_float4(0, 0, 0, 0), m4 = p4, v4 = p4;
    if (VEC && active) {
      p4 = ld_f4(P + off); m4 = ld_f4(Mo + off); v4 = ld_f4(Vo + off);
    }
    float g[4];
    segment_sum<VEC>(gs, sorted, perm, i, R, row, col, active, g);
    if (!active) continue;
    if (VEC) {
      float pp[4] = {p4.x, p4.y, p4.z, p4.w}, mm[4] = {m4.x, m4.y, m4.z, m4.w}, vv[4] = {v4.x, v4.y, v4.z, v4.w};
#pragma unroll
      for (int e = 0; e < 4; ++e) sparse_adam_elem(g[e], pp[e], mm[e], vv[e], s);
      st_f4(P + off, make_float4(pp[0], pp[1], pp[2], pp[3]));
      st_f4(Mo + off, make_float4(mm[0], mm[1], mm[2], mm[3]));
      st_f4(Vo + off, make_float4(vv[0], vv[1], vv[2], vv[3]));
    } else {
      sparse_adam_update<VEC>(P, Mo, Vo, off, g, s);
    }
  }
}

template <bool VEC>
__global__ void __launch_bounds__(32 * kLongWarps) sparse_adam_long_kernel(float* __restrict__ P, float* __restrict__ Mo,
                                                               float* __restrict__ Vo, int D,
                                                               const int64_t* __restrict__ sorted,
                                                               const int32_t* __restrict__ perm, int64_t R, GradSrc gs,
                                                               AdamScalars s, const float* __restrict__ scalars,
                                                               const ttam_step_state* __restrict__ st,
                                                               const int32_t* __restrict__ list) {
  __shared__ float part[kLongWarps][128];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (st) s.step_size = scalars[4 * st->step + 2];
  const int n = list[0];
  constexpr int W = VEC ? 128 : 32;
  for (int e = blockIdx.x; e < n; e += gridDim.x) {
    const int64_t i = list[1 + e];
    const int64_t row = sorted[i];
    const int64_t end = segment_end(sorted, i, R, row, lane);
    for (int c0 = 0; c0 < D; c0 += W) {
      const int col = c0 + (VEC ? lane * 4 : lane);
      const bool active = col < D;
      float g[4];
      segment_sum_block<VEC>(gs, perm, i, end, col, active, part, g);
      if (warp == 0 && active) sparse_adam_update<VEC>(P, Mo, Vo, row * (int64_t)D + col, g, s);
    }
  }
}

// Replay zero-gradient steps (t_from, t_to] ... i.e. steps t_from+1 .. t_to, on NE elements held in registers.
//
// A row that is touched every ~50 (items) / ~120 (users) steps replays that many steps each time, so this loop is most of
// what the lazily-updated tables cost once training is under way (measured with dense_elem in the loop - IEEE sqrt and two
// IEEE divisions, 52 instructions per element-step: the train step grew from 1.3 ms at step 10 to 2.3 ms at step 1000).
// With a zero gradient (AdamW, or Adam without weight decay) the recurrence is
//     p <- p*(1 - lr*wd);  m <- m - (1-b1)*m;  v <- b2*v;  p <- p - (lr/bc1_t) * m / (sqrt(v)/sqrt(bc2_t) + eps)
// so v at step t is v0 * b2^(t - t0) and sqrt(v_t)/sqrt(bc2_t) = [sqrt(v0) / b2^(t0/2)] * h_t with the per-step scalar
// h_t = b2^(t/2) / sqrt(1 - b2^t) (scalar table [4t+3], computed in double from the fp32-rounded b2 the reference multiplies
// by).  The fast path (default) evaluates exactly that: one FMA for the denominator, MUFU.RCP + one multiply for the
// division (2 ulp), no square root and no recurrence on v inside the loop - 6 instructions per element-step - and writes
// v0 * (b2^(t1/2) / b2^(t0/2))^2 back.  p and m follow the reference's own fp32 recurrences.  Versus the step-by-step
// arithmetic each replayed step differs by a few ulp of its UPDATE term (|update| <= ~lr, i.e. ~1e-10 absolute on
// parameters of magnitude 1e-2); tests/test_gpu_kernels.py bounds the drift over 1000 replayed steps against the
// reference's step-by-step fp32 arithmetic.
// m == 0 (a row no gradient has reached yet) makes every update term exactly zero: only the decay is applied.
// TTAM_EXACT_REPLAY=1 selects the step-by-step loop.
template <int KIND, int NE>
__device__ __forceinline__ void replay(float (&pp)[NE], float (&mm)[NE], float (&vv)[NE], int t_from, int t_to,
                                       const float* __restrict__ scalars, const AdamScalars& s) {
  if (t_from >= t_to) return;
  const bool zero_grad = KIND == TTAM_OPT_ADAMW || (KIND == TTAM_OPT_ADAM && s.wd == 0.f);
  // b2^(t/2) = h_t * sqrt(bc2_t); t = 0 -> 1
  const float g_from = t_from > 0 ? __fmul_rn(scalars[4 * t_from + 3], scalars[4 * t_from + 1]) : 1.f;
  if (zero_grad && s.fast_replay && g_from > 1e-30f) {
    const bool decay = KIND == TTAM_OPT_ADAMW && s.wd != 0.f;
    const float ratio = __fdiv_rn(__fmul_rn(scalars[4 * t_to + 3], scalars[4 * t_to + 1]), g_from);   // b2^((t1-t0)/2)
    bool m_zero = true;
#pragma unroll
    for (int e = 0; e < NE; ++e) m_zero = m_zero && (mm[e] == 0.f);
    if (m_zero) {
      if (decay) {
        for (int t = t_from + 1; t <= t_to; ++t) {
#pragma unroll
          for (int e = 0; e < NE; ++e) pp[e] = __fmul_rn(pp[e], s.decay);
        }
      }
    } else {
      float w0[NE];
#pragma unroll
      for (int e = 0; e < NE; ++e) w0[e] = __fdiv_rn(sqrtf(vv[e]), g_from);
      for (int t = t_from + 1; t <= t_to; ++t) {
        const float a = scalars[4 * t], h = scalars[4 * t + 3];
#pragma unroll
        for (int e = 0; e < NE; ++e) {
          if (decay) pp[e] = __fmul_rn(pp[e], s.decay);
          mm[e] = fmaf(s.one_minus_b1, -mm[e], mm[e]);
          pp[e] = fmaf(-a, __fdividef(mm[e], fmaf(w0[e], h, s.eps)), pp[e]);
        }
      }
    }
#pragma unroll
    for (int e = 0; e < NE; ++e) vv[e] = __fmul_rn(__fmul_rn(vv[e], ratio), ratio);
    return;
  }
  for (int t = t_from + 1; t <= t_to; ++t) {
    const float step_size = KIND == TTAM_OPT_SGD ? 0.f : scalars[4 * t];
    const float bc2s = KIND == TTAM_OPT_SGD ? 1.f : scalars[4 * t + 1];
#pragma unroll
    for (int e = 0; e < NE; ++e) dense_elem<KIND>(0.f, pp[e], mm[e], vv[e], s, step_size, bc2s);
  }
}

template <int KIND, bool VEC>
__device__ __forceinline__ void lazy_update(float* __restrict__ P, float* __restrict__ Mo, float* __restrict__ Vo,
                                            int64_t off, const float (&g)[4], int t_prev, int step,
                                            const float* __restrict__ scalars, const AdamScalars& s) {
  const bool has_v = KIND != TTAM_OPT_SGD;
  const bool has_m = KIND != TTAM_OPT_SGD || s.momentum != 0.f;
  constexpr int NE = VEC ? 4 : 1;
  const float step_size = KIND == TTAM_OPT_SGD ? 0.f : scalars[4 * step];
  const float bc2s = KIND == TTAM_OPT_SGD ? 1.f : scalars[4 * step + 1];
  float pp[NE], mm[NE], vv[NE];
  if (VEC) {
    float4 p4 = ld_f4(P + off);
    float4 m4 = has_m ? ld_f4(Mo + off) : make_float4(0, 0, 0, 0);
    float4 v4 = has_v ? ld_f4(Vo + off) : make_float4(0, 0, 0, 0);
    pp[0] = p4.x; mm[0] = m4.x; vv[0] = v4.x;
    if (NE == 4) {
      pp[1 % NE] = p4.y; pp[2 % NE] = p4.z; pp[3 % NE] = p4.w;
      mm[1 % NE] = m4.y; mm[2 % NE] = m4.z; mm[3 % NE] = m4.w;
      vv[1 % NE] = v4.y; vv[2 % NE] = v4.z; vv[3 % NE] = v4.w;
    }
  } else {
    pp[0] = P[off];
    mm[0] = has_m ? Mo[off] : 0.f;
    vv[0] = has_v ? Vo[off] : 0.f;
  }
  replay<KIND, NE>(pp, mm, vv, t_prev, step - 1, scalars, s);
#pragma unroll
  for (int e = 0; e < NE; ++e) dense_elem<KIND>(g[e], pp[e], mm[e], vv[e], s, step_size, bc2s);
  if (VEC) {
    st_f4(P + off, make_float4(pp[0], pp[1 % NE], pp[2 % NE], pp[3 % NE]));
    if (has_m) st_f4(Mo + off, make_float4(mm[0], mm[1 % NE], mm[2 % NE], mm[3 % NE]));
    if (has_v) st_f4(Vo + off, make_float4(vv[0], vv[1 % NE], vv[2 % NE], vv[3 % NE]));
  } else {
    P[off] = pp[0];
    if (has_m) Mo[off] = mm[0];
    if (has_v) Vo[off] = vv[0];
  }
}

template <int KIND, bool VEC>
__global__ void __launch_bounds__(256, 6) lazy_rows_kernel(float* __restrict__ P, float* __restrict__ Mo,
                                                        float* __restrict__ Vo, int32_t* __restrict__ last_step, int D,
                                                        const int64_t* __restrict__ sorted,
                                                        const int32_t* __restrict__ perm, int64_t R, GradSrc gs,
                                                        const float* __restrict__ scalars, AdamScalars s, int step,
                                                        const ttam_step_state* __restrict__ st, bool skip_long) {
  const int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (i >= R) return;
  const int64_t row = sorted[i];
  if (i > 0 && sorted[i - 1] == row) return;
  if (skip_long && is_long_segment(sorted, i, R, row)) return;
  if (st) step = st->step;
  const int t_prev = last_step[row];
  constexpr int W = VEC ? 128 : 32;
  for (int c0 = 0; c0 < D; c0 += W) {
    const int col = c0 + (VEC ? lane * 4 : lane);
    const bool active = col < D;
    if (active) {  // start fetching the row's p, m, v while the gradient rows are summed (independent latency chains)
      const int64_t off = row * (int64_t)D + col;
      asm volatile("prefetch.global.L2 [%0];" ::"l"(P + off));
      if (Mo) asm volatile("prefetch.global.L2 [%0];" ::"l"(Mo + off));
      if (Vo) asm volatile("prefetch.global.L2 [%0];" ::"l"(Vo + off));
    }
    float g[4];
    segment_sum<VEC>(gs, sorted, perm, i, R, row, col, active, g);
    if (!active) continue;
    lazy_update<KIND, VEC>(P, Mo, Vo, row * (int64_t)D + col, g, t_prev, step, scalars, s);
  }
  __syncwarp();
  if (lane == 0) last_step[row] = step;
}

template <int KIND, bool VEC>
__global__ void __launch_bounds__(32 * kLongWarps) lazy_long_kernel(float* __restrict__ P, float* __restrict__ Mo,
                                                        float* __restrict__ Vo, int32_t* __restrict__ last_step, int D,
                                                        const int64_t* __restrict__ sorted,
                                                        const int32_t* __restrict__ perm, int64_t R, GradSrc gs,
                                                        const float* __restrict__ scalars, AdamScalars s, int step,
                                                        const ttam_step_state* __restrict__ st,
                                                        const int32_t* __restrict__ list) {
  __shared__ float part[kLongWarps][128];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (st) step = st->step;
  const int n = list[0];
  constexpr int W = VEC ? 128 : 32;
  for (int e = blockIdx.x; e < n; e += gridDim.x) {
    const int64_t i = list[1 + e];
    const int64_t row = sorted[i];
    const int64_t end = segment_end(sorted, i, R, row, lane);
    const int t_prev = last_step[row];
    for (int c0 = 0; c0 < D; c0 += W) {
      const int col = c0 + (VEC ? lane * 4 : lane);
      const bool active = col < D;
      float g[4];
      segment_sum_block<VEC>(gs, perm, i, end, col, active, part, g);
      if (warp == 0 && active) lazy_update<KIND, VEC>(P, Mo, Vo, row * (int64_t)D + col, g, t_prev, step, scalars, s);
    }
    __syncthreads();  // every warp has read last_step[row] before it changes
    if (threadIdx.x == 0) last_step[row] = step;
  }
}

// Bring the unique rows of `sorted` up to step-1 (zero-gradient replay) so that the forward pass of step `step`
// reads current values; the row-wise update at the end of the step then finds nothing left to replay.
template <int KIND, bool VEC>
__global__ void __launch_bounds__(256, 6) lazy_catchup_kernel(float* __restrict__ P, float* __restrict__ Mo,
                                                           float* __restrict__ Vo, int32_t* __restrict__ last_step, int D,
                                                           const int64_t* __restrict__ sorted, int64_t R,
                                                           const float* __restrict__ scalars, AdamScalars s, int step,
                                                           const ttam_step_state* __restrict__ st) {
  const int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (i >= R) return;
  const int64_t row = sorted[i];
  if (i > 0 && sorted[i - 1] == row) return;
  if (st) step = st->step;
  const int t_prev = last_step[row];
  if (t_prev >= step - 1) return;
  const bool has_v = KIND != TTAM_OPT_SGD;
  const bool has_m = KIND != TTAM_OPT_SGD || s.momentum != 0.f;
  constexpr int W = VEC ? 128 : 32;
  constexpr int NE = VEC ? 4 : 1;
  for (int c0 = 0; c0 < D; c0 += W) {
    const int col = c0 + (VEC ? lane * 4 : lane);
    if (col >= D) continue;
    const int64_t off = row * (int64_t)D + col;
    float pp[NE], mm[NE], vv[NE];
    if (VEC) {
      float4 p4 = ld_f4(P + off);
      float4 m4 = has_m ? ld_f4(Mo + off) : make_float4(0, 0, 0, 0);
      float4 v4 = has_v ? ld_f4(Vo + off) : make_float4(0, 0, 0, 0);
      pp[0] = p4.x; mm[0] = m4.x; vv[0] = v4.x;
      if (NE == 4) {
        pp[1 % NE] = p4.y; pp[2 % NE] = p4.z; pp[3 % NE] = p4.w;
        mm[1 % NE] = m4.y; mm[2 % NE] = m4.z; mm[3 % NE] = m4.w;
        vv[1 % NE] = v4.y; vv[2 % NE] = v4.z; vv[3 % NE] = v4.w;
      }
    } else {
      pp[0] = P[off];
      mm[0] = has_m ? Mo[off] : 0.f;
      vv[0] = has_v ? Vo[off] : 0.f;
    }
    replay<KIND, NE>(pp, mm, vv, t_prev, step - 1, scalars, s);
    if (VEC) {
      st_f4(P + off, make_float4(pp[0], pp[1 % NE], pp[2 % NE], pp[3 % NE]));
      if (has_m) st_f4(Mo + off, make_float4(mm[0], mm[1 % NE], mm[2 % NE], mm[3 % NE]));
      if (has_v) st_f4(Vo + off, make_float4(vv[0], vv[1 % NE], vv[2 % NE], vv[3 % NE]));
    } else {
      P[off] = pp[0];
      if (has_m) Mo[off] = mm[0];
      if (has_v) Vo[off] = vv[0];
    }
  }
  __syncwarp();
  if (lane == 0) last_step[row] = step - 1;
}

template <int KIND, bool VEC>
__global__ void __launch_bounds__(256) lazy_flush_kernel(float* __restrict__ P, float* __restrict__ Mo,
                                                         float* __restrict__ Vo, int32_t* __restrict__ last_step,
                                                         int64_t num_rows, int D, const float* __restrict__ scalars,
                                                         AdamScalars s, int step, const ttam_step_state* __restrict__ st) {
  if (st) step = st->step;
  constexpr int NE = VEC ? 4 : 1;
  const int cpr = VEC ? D / 4 : D;  // work items per row
  const int64_t total = num_rows * cpr;
  const bool has_v = KIND != TTAM_OPT_SGD;
  const bool has_m = KIND != TTAM_OPT_SGD || s.momentum != 0.f;
  for (int64_t w = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; w < total; w += (int64_t)gridDim.x * blockDim.x) {
    const int64_t row = w / cpr;
    const int c = (int)(w - row * cpr);
    const int t_prev = last_step[row];
    if (t_prev >= step) continue;
    const int64_t off = row * (int64_t)D + (VEC ? c * 4 : c);
    float pp[NE], mm[NE], vv[NE];
    if (VEC) {
      float4 p4 = ld_f4(P + off);
      float4 m4 = has_m ? ld_f4(Mo + off) : make_float4(0, 0, 0, 0);
      float4 v4 = has_v ? ld_f4(Vo + off) : make_float4(0, 0, 0, 0);
      pp[0] = p4.x; mm[0] = m4.x; vv[0] = v4.x;
      if (NE == 4) {
        pp[1 % NE] = p4.y; pp[2 % NE] = p4.z; pp[3 % NE] = p4.w;
        mm[1 % NE] = m4.y; mm[2 % NE] = m4.z; mm[3 % NE] = m4.w;
        vv[1 % NE] = v4.y; vv[2 % NE] = v4.z; vv[3 % NE] = v4.w;
      }
    } else {
      pp[0] = P[off];
      mm[0] = has_m ? Mo[off] : 0.f;
      vv[0] = has_v ? Vo[off] : 0.f;
    }
    replay<KIND, NE>(pp, mm, vv, t_prev, step, scalars, s);
    if (VEC) {
      st_f4(P + off, make_float4(pp[0], pp[1 % NE], pp[2 % NE], pp[3 % NE]));
      if (has_m) st_f4(Mo + off, make_float4(mm[0], mm[1 % NE], mm[2 % NE], mm[3 % NE]));
      if (has_v) st_f4(Vo + off, make_float4(vv[0], vv[1 % NE], vv[2 % NE], vv[3 % NE]));
    } else {
      P[off] = pp[0];
      if (has_m) Mo[off] = mm[0];
      if (has_v) Vo[off] = vv[0];
    }
  }
}

// second pass of the flush: stamp the rows (separate so that all chunks of a row saw the old stamp)
__global__ void stamp_kernel(int32_t* last_step, int64_t num_rows, int step, const ttam_step_state* __restrict__ st) {
  if (st) step = st->step;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < num_rows; i += (int64_t)gridDim.x * blockDim.x)
    if (last_step[i] < step) last_step[i] = step;
}

template <int KIND>
__global__ void __launch_bounds__(256) dense_step_kernel(ttam_tensor_list list, AdamScalars s, float step_size,
                                                         float bc2s, const float* __restrict__ scalars,
                                                         const ttam_step_state* __restrict__ st) {
  if (st && KIND != TTAM_OPT_SGD) {
    step_size = scalars[4 * st->step];
    bc2s = scalars[4 * st->step + 1];
  }
  const int ti = blockIdx.y;
  float* __restrict__ p = list.p[ti];
  const float* __restrict__ g = list.g[ti];
  float* __restrict__ m = list.m[ti];
  float* __restrict__ v = list.v[ti];
  const int64_t n = list.numel[ti];
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    float pp = p[i], mm = m ? m[i] : 0.f, vv = v ? v[i] : 0.f;
    dense_elem<KIND>(g[i], pp, mm, vv, s, step_size, bc2s);
    p[i] = pp;
    if (m) m[i] = mm;
    if (v) v[i] = vv;
  }
}

// Hyper-parameters arrive as doubles (Python floats) and are rounded to fp32 exactly where torch rounds them:
// each scalar expression is evaluated in double first (`1 - beta1`, `1 - lr * weight_decay`, ...).
static AdamScalars make_scalars(double lr, double wd, double b1, double b2, double eps, double momentum) {
  AdamScalars s{};
  s.lr = (float)lr; s.wd = (float)wd; s.b1 = (float)b1; s.b2 = (float)b2; s.eps = (float)eps; s.momentum = (float)momentum;
  s.one_minus_b1 = (float)(1.0 - b1);
  s.one_minus_b2 = (float)(1.0 - b2);
  s.decay = (float)(1.0 - lr * wd);
  static const int exact = [] { const char* e = getenv("TTAM_EXACT_REPLAY"); return (e && e[0] == '1') ? 1 : 0; }();
  s.fast_replay = exact ? 0 : 1;
  return s;
}

}  // namespace ttam

using namespace ttam;

extern "C" int64_t ttam_sort_workspace_bytes(int64_t R) {
  if (R <= 0) return 256;
  size_t a = sort_temp_bytes(R), b = unique_temp_bytes(R);
  return (int64_t)align_up((int64_t)(a > b ? a : b), 256) + align_up(R * 4, 256) + 256;
}

extern "C" int ttam_sort_rows(const int64_t* idx, int64_t R, int64_t num_rows, int64_t* sorted_idx, int32_t* perm,
                              void* workspace, int64_t workspace_bytes, void* stream) {
  TTAM_CHECK_ARG(idx && sorted_idx && perm && workspace, "sort_rows: null pointer");
  TTAM_CHECK_ARG(R >= 0 && R < (1ll << 31) && num_rows > 0, "sort_rows: bad size");
  if (R == 0) return TTAM_OK;
  if (workspace_bytes < ttam_sort_workspace_bytes(R)) {
    set_error("sort_rows: workspace too small");
    return TTAM_EWORKSPACE;
  }
  cudaStream_t s = (cudaStream_t)stream;
  int32_t* iota = (int32_t*)workspace;
  void* temp = (char*)workspace + align_up(R * 4, 256);
  size_t temp_bytes = sort_temp_bytes(R);
  iota_kernel<<<(int)std::min<int64_t>(ceil_div(R, 256), 1024), 256, 0, s>>>(iota, R);
  TTAM_LAUNCH_CHECK();
  TTAM_CUDA(cub::DeviceRadixSort::SortPairs(temp, temp_bytes, idx, sorted_idx, (const int32_t*)iota, perm, (int)R, 0,
                                            bits_for(num_rows), s));
  return TTAM_OK;
}

extern "C" int ttam_unique_rows(const int64_t* sorted_idx, int64_t R, int64_t* unique_out, int64_t* n_unique_out,
                                void* workspace, int64_t workspace_bytes, void* stream) {
  TTAM_CHECK_ARG(sorted_idx && unique_out && n_unique_out && workspace, "unique_rows: null pointer");
  if (workspace_bytes < ttam_sort_workspace_bytes(R)) {
    set_error("unique_rows: workspace too small");
    return TTAM_EWORKSPACE;
  }
  size_t temp_bytes = unique_temp_bytes(R);
  TTAM_CUDA(cub::DeviceSelect::Unique(workspace, temp_bytes, sorted_idx, unique_out, n_unique_out, (int)R,
                                      (cudaStream_t)stream));
  return TTAM_OK;
}

extern "C" int64_t ttam_long_segments_bytes(int64_t R) { return align_up((R / kLongSeg + 2) * 4, 256); }

extern "C" int ttam_find_long_segments(const int64_t* sorted_idx, int64_t R, int32_t* long_list, void* stream) {
  TTAM_CHECK_ARG(long_list && (R == 0 || sorted_idx), "find_long_segments: null pointer");
  cudaStream_t s = (cudaStream_t)stream;
  TTAM_CUDA(cudaMemsetAsync(long_list, 0, 4, s));
  if (R <= kLongSeg) return TTAM_OK;
  find_long_segments_kernel<<<(int)ceil_div(R, 256), 256, 0, s>>>(sorted_idx, R, long_list);
  TTAM_LAUNCH_CHECK();
  return TTAM_OK;
}

static bool vec_ok(int64_t D, const void* p, const void* m, const void* v, const float* ga, int64_t lda,
                   const float* gb, int64_t ldb) {
  auto al = [](const void* q) { return q == nullptr || ((uintptr_t)q & 15) == 0; };
  return D % 4 == 0 && al(p) && al(m) && al(v) && al(ga) && al(gb) && lda % 4 == 0 && (gb == nullptr || ldb % 4 == 0);
}

extern "C" int ttam_sparse_adam_rows(float* p, float* m, float* v, int64_t D, const int64_t* sorted_idx,
                                     const int32_t* perm, int64_t R, const float* grad_a, int64_t ld_a, int64_t n_a,
                                     const float* grad_b, int64_t ld_b, const float* scalars, double lr, double beta1,
                                     double beta2, double eps, int64_t step, const ttam_step_state* state_dev,
                                     const int32_t* long_list, void* stream) {
  TTAM_CHECK_ARG(p && m && v && sorted_idx && perm && grad_a, "sparse_adam_rows: null pointer");
  TTAM_CHECK_ARG(!state_dev || scalars, "sparse_adam_rows: a device step needs the scalar table");
  TTAM_CHECK_ARG(D > 0 && step >= 1 && n_a >= 0 && n_a <= R && (n_a == R || grad_b), "sparse_adam_rows: bad argument");
  if (R == 0) return TTAM_OK;
  AdamScalars s = make_scalars(lr, 0.0, beta1, beta2, eps, 0.0);
  const double bc1 = 1.0 - pow(beta1, (double)step), bc2 = 1.0 - pow(beta2, (double)step);
  s.step_size = (float)(lr * sqrt(bc2) / bc1);
  GradSrc gs{grad_a, ld_a, n_a, grad_b, ld_b};
  const int blocks = (int)ceil_div(R * 32, 256);
  const bool vec = vec_ok(D, p, m, v, grad_a, ld_a, grad_b, ld_b);
  const bool skip = long_list != nullptr;
  cudaStream_t cs = (cudaStream_t)stream;
  if (vec) sparse_adam_rows_kernel<true><<<blocks, 256, 0, cs>>>(p, m, v, (int)D, sorted_idx, perm, R, gs, s, scalars, state_dev, skip);
  else sparse_adam_rows_kernel<false><<<blocks, 256, 0, cs>>>(p, m, v, (int)D, sorted_idx, perm, R, gs, s, scalars, state_dev, skip);
  TTAM_LAUNCH_CHECK();
  if (skip) {
    if (vec) sparse_adam_long_kernel<true><<<2 * num_sms(), 32 * kLongWarps, 0, cs>>>(p, m, v, (int)D, sorted_idx, perm, R, gs, s, scalars, state_dev, long_list);
    else sparse_adam_long_kernel<false><<<2 * num_sms(), 32 * kLongWarps, 0, cs>>>(p, m, v, (int)D, sorted_idx, perm, R, gs, s, scalars, state_dev, long_list);
    TTAM_LAUNCH_CHECK();
  }
  return TTAM_OK;
}

#define TTAM_DISPATCH_KIND(kind, VEC, CALL)                                   \
  do {                                                                        \
    if (kind == TTAM_OPT_ADAMW) { CALL(TTAM_OPT_ADAMW, VEC); }                \
    else if (kind == TTAM_OPT_ADAM) { CALL(TTAM_OPT_ADAM, VEC); }             \
    else { CALL(TTAM_OPT_SGD, VEC); }                                         \
  } while (0)

extern "C" int ttam_lazy_rows(int kind, float* p, float* m, float* v, int32_t* last_step, int64_t D,
                              const int64_t* sorted_idx, const int32_t* perm, int64_t R, const float* grad_a,
                              int64_t ld_a, int64_t n_a, const float* grad_b, int64_t ld_b, const float* scalars,
                              double lr, double weight_decay, double beta1, double beta2, double eps, double momentum,
                              int64_t step, const ttam_step_state* state_dev, const int32_t* long_list, void* stream) {
  TTAM_CHECK_ARG(p && last_step && sorted_idx && perm && grad_a, "lazy_rows: null pointer");
  TTAM_CHECK_ARG(kind >= TTAM_OPT_ADAMW && kind <= TTAM_OPT_SGD, "lazy_rows: unknown optimiser kind %d", kind);
  TTAM_CHECK_ARG(kind == TTAM_OPT_SGD || (m && v && scalars), "lazy_rows: Adam needs m, v and the scalar table");
  TTAM_CHECK_ARG(kind != TTAM_OPT_SGD || momentum == 0.0 || m, "lazy_rows: SGD momentum needs the buffer m");
  TTAM_CHECK_ARG(D > 0 && step >= 1 && step < (1ll << 30) && n_a >= 0 && n_a <= R && (n_a == R || grad_b),
                 "lazy_rows: bad argument");
  if (R == 0) return TTAM_OK;
  AdamScalars s = make_scalars(lr, weight_decay, beta1, beta2, eps, momentum);
  GradSrc gs{grad_a, ld_a, n_a, grad_b, ld_b};
  const int blocks = (int)ceil_div(R * 32, 256);
  cudaStream_t st = (cudaStream_t)stream;
  const bool skip = long_list != nullptr;
  const bool vec = vec_ok(D, p, m, v, grad_a, ld_a, grad_b, ld_b);
#define CALL(K, V) lazy_rows_kernel<K, V><<<blocks, 256, 0, st>>>(p, m, v, last_step, (int)D, sorted_idx, perm, R, gs, scalars, s, (int)step, state_dev, skip)
  if (vec) TTAM_DISPATCH_KIND(kind, true, CALL);
  else TTAM_DISPATCH_KIND(kind, false, CALL);
#undef CALL
  TTAM_LAUNCH_CHECK();
  if (skip) {
#define CALL(K, V) lazy_long_kernel<K, V><<<2 * num_sms(), 32 * kLongWarps, 0, st>>>(p, m, v, last_step, (int)D, sorted_idx, perm, R, gs, scalars, s, (int)step, state_dev, long_list)
    if (vec) TTAM_DISPATCH_KIND(kind, true, CALL);
    else TTAM_DISPATCH_KIND(kind, false, CALL);
#undef CALL
    TTAM_LAUNCH_CHECK();
  }
  return TTAM_OK;
}

extern "C" int ttam_lazy_catchup(int kind, float* p, float* m, float* v, int32_t* last_step, int64_t D,
                                 const int64_t* sorted_idx, int64_t R, const float* scalars, double lr,
                                 double weight_decay, double beta1, double beta2, double eps, double momentum,
                                 int64_t step, const ttam_step_state* state_dev, void* stream) {
  TTAM_CHECK_ARG(kind >= TTAM_OPT_ADAMW && kind <= TTAM_OPT_SGD, "lazy_catchup: unknown optimiser kind %d", kind);
  if (R == 0) return TTAM_OK;
  TTAM_CHECK_ARG(p && last_step && sorted_idx, "lazy_catchup: null pointer");
  TTAM_CHECK_ARG(kind == TTAM_OPT_SGD || (m && v && scalars), "lazy_catchup: Adam needs m, v and the scalar table");
  TTAM_CHECK_ARG(kind != TTAM_OPT_SGD || momentum == 0.0 || m, "lazy_catchup: SGD momentum needs the buffer m");
  TTAM_CHECK_ARG(D > 0 && step >= 1 && step < (1ll << 30), "lazy_catchup: bad argument");
  AdamScalars s = make_scalars(lr, weight_decay, beta1, beta2, eps, momentum);
  const int blocks = (int)ceil_div(R * 32, 256);
  cudaStream_t st = (cudaStream_t)stream;
#define CALL(K, V) lazy_catchup_kernel<K, V><<<blocks, 256, 0, st>>>(p, m, v, last_step, (int)D, sorted_idx, R, scalars, s, (int)step, state_dev)
  if (vec_ok(D, p, m, v, nullptr, 0, nullptr, 0)) TTAM_DISPATCH_KIND(kind, true, CALL);
  else TTAM_DISPATCH_KIND(kind, false, CALL);
#undef CALL
  TTAM_LAUNCH_CHECK();
  return TTAM_OK;
}

extern "C" int ttam_lazy_flush(int kind, float* p, float* m, float* v, int32_t* last_step, int64_t num_rows, int64_t D,
                               const float* scalars, double lr, double weight_decay, double beta1, double beta2, double eps,
                               double momentum, int64_t step, const ttam_step_state* state_dev, void* stream) {
  TTAM_CHECK_ARG(p && last_step, "lazy_flush: null pointer");
  TTAM_CHECK_ARG(kind >= TTAM_OPT_ADAMW && kind <= TTAM_OPT_SGD, "lazy_flush: unknown optimiser kind %d", kind);
  TTAM_CHECK_ARG(kind == TTAM_OPT_SGD || (m && v && scalars), "lazy_flush: Adam needs m, v and the scalar table");
  TTAM_CHECK_ARG(D > 0 && step >= 0 && num_rows >= 0, "lazy_flush: bad argument");
  if (num_rows == 0 || (step == 0 && !state_dev)) return TTAM_OK;
  AdamScalars s = make_scalars(lr, weight_decay, beta1, beta2, eps, momentum);
  cudaStream_t st = (cudaStream_t)stream;
  const bool vec = vec_ok(D, p, m, v, nullptr, 0, nullptr, 0);
  const int64_t total = num_rows * (vec ? D / 4 : D);
  const int blocks = (int)std::min<int64_t>(ceil_div(total, 256), (int64_t)num_sms() * 32);
#define CALL(K, V) lazy_flush_kernel<K, V><<<blocks, 256, 0, st>>>(p, m, v, last_step, num_rows, (int)D, scalars, s, (int)step, state_dev)
  if (vec) TTAM_DISPATCH_KIND(kind, true, CALL);
  else TTAM_DISPATCH_KIND(kind, false, CALL);
#undef CALL
  TTAM_LAUNCH_CHECK();
  stamp_kernel<<<(int)std::min<int64_t>(ceil_div(num_rows, 256), 4096), 256, 0, st>>>(last_step, num_rows, (int)step, state_dev);
  TTAM_LAUNCH_CHECK();
  return TTAM_OK;
}

extern "C" int ttam_dense_step(int kind, const ttam_tensor_list* list_host, const float* scalars, double lr,
                               double weight_decay, double beta1, double beta2, double eps, double momentum, int64_t step,
                               const ttam_step_state* state_dev, void* stream) {
  TTAM_CHECK_ARG(list_host, "dense_step: null list");
  TTAM_CHECK_ARG(!state_dev || scalars || kind == TTAM_OPT_SGD, "dense_step: a device step needs the scalar table");
  TTAM_CHECK_ARG(list_host->count >= 0 && list_host->count <= TTAM_MAX_TENSORS, "dense_step: too many tensors");
  TTAM_CHECK_ARG(kind >= TTAM_OPT_ADAMW && kind <= TTAM_OPT_SGD && step >= 1, "dense_step: bad argument");
  if (list_host->count == 0) return TTAM_OK;
  int64_t max_n = 0;
  for (int i = 0; i < list_host->count; ++i) {
    TTAM_CHECK_ARG(list_host->p[i] && list_host->g[i], "dense_step: tensor %d has a null pointer", i);
    TTAM_CHECK_ARG(kind == TTAM_OPT_SGD || (list_host->m[i] && list_host->v[i]), "dense_step: Adam needs m and v");
    if (list_host->numel[i] > max_n) max_n = list_host->numel[i];
  }
  AdamScalars s = make_scalars(lr, weight_decay, beta1, beta2, eps, momentum);
  const double bc1 = 1.0 - pow(beta1, (double)step), bc2 = 1.0 - pow(beta2, (double)step);
  const float step_size = (float)(lr / bc1), bc2s = (float)sqrt(bc2);
  dim3 grid((unsigned)std::min<int64_t>(ceil_div(max_n, 256), 512), (unsigned)list_host->count, 1);
  cudaStream_t st = (cudaStream_t)stream;
  if (kind == TTAM_OPT_ADAMW) dense_step_kernel<TTAM_OPT_ADAMW><<<grid, 256, 0, st>>>(*list_host, s, step_size, bc2s, scalars, state_dev);
  else if (kind == TTAM_OPT_ADAM) dense_step_kernel<TTAM_OPT_ADAM><<<grid, 256, 0, st>>>(*list_host, s, step_size, bc2s, scalars, state_dev);
  else dense_step_kernel<TTAM_OPT_SGD><<<grid, 256, 0, st>>>(*list_host, s, step_size, bc2s, scalars, state_dev);
  TTAM_LAUNCH_CHECK();
  return TTAM_OK;
}
