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

// ---- row-wise update kernels ---------------------------------------------------------------------------------------
// HBM-bound: per unique row one read + one write of p, m, v (2304 B at D = 96) plus the row's gradient rows.
//
// Window role.  One warp owns a WINDOW of 32 consecutive positions of the sorted index list.  Its index work is
// lane-parallel and happens once: lane l loads sorted[base + l] and perm[base + l] (lanes 0..15 also the 16 positions
// behind the window, so that a segment that starts inside the window can be followed across its end), segment heads
// come out of one shuffle + ballot.  The dependent chain  sorted -> perm -> gradient row  of the one-warp-per-position
// kernels this replaces (three latencies for every 384-byte row) is paid once per window, and the rows of kHB segment
// heads are fetched together: 4 heads x (p, m, v, first gradient row) = 16 independent 16-byte loads per lane in
// flight before the first one is used.  Further members of a segment are added in sorted order (deterministic).
//
// Long role.  A popular row (Zipf-distributed positives) can own hundreds of gradient rows of a batch; segments longer
// than kLongSeg are listed by find_long_segments_kernel and summed by a whole block each - every warp takes a strided
// share with kLongUnroll loads in flight, its perm entries fetched 32 at a time by the lanes, the partial sums combined
// in warp order (a fixed order: deterministic).  The first `long_blocks` blocks of the SAME launch take that role, so
// the long segments run next to the windows instead of behind them.
constexpr int kLongSeg = 16;
constexpr int kRowThreads = 256;
constexpr int kRowWarps = kRowThreads / 32;
constexpr int kCatchHB = 4;      // rows in flight per warp (catch-up)
constexpr int kCatchStride = 8;  // sorted positions per warp (catch-up: arithmetic-bound, many short warps balance better)
constexpr int kLongUnroll = 8;   // gradient rows in flight per warp (long role)
constexpr unsigned kFull = 0xffffffffu;

__global__ void __launch_bounds__(256) find_long_segments_kernel(const int64_t* __restrict__ sorted, int64_t R,
                                                                  int32_t* __restrict__ list) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i + kLongSeg >= R) return;
  const int64_t row = sorted[i];
  if (i > 0 && sorted[i - 1] == row) return;
  if (sorted[i + kLongSeg] == row) list[1 + atomicAdd(list, 1)] = (int32_t)i;
}

// first position after the segment that starts at i (32-ary search by one warp; every lane returns the result)
__device__ __forceinline__ int64_t segment_end(const int64_t* __restrict__ sorted, int64_t i, int64_t R, int64_t row, int lane) {
  int64_t lo = i, hi = R;  // sorted[lo] == row, (hi == R or sorted[hi] != row)
  while (hi - lo > 1) {
    const int64_t span = hi - lo;
    const int64_t stepw = (span + 31) / 32;
    const int64_t probe = lo + (int64_t)(lane + 1) * stepw;
    const bool same = probe < hi && sorted[probe] == row;
    const unsigned bal = __ballot_sync(kFull, same);
    const int k = __popc(bal);  // probes 1..k hit the row (the predicate is monotone)
    const int64_t new_lo = lo + (int64_t)k * stepw;
    const int64_t new_hi = lo + (int64_t)(k + 1) * stepw;
    lo = new_lo;
    hi = new_hi < hi ? new_hi : hi;
  }
  return hi;
}

template <bool VEC>
__device__ __forceinline__ void ld_elems(const float* __restrict__ src, float (&v)[4]) {
  if (VEC) {
    const float4 f = ld_f4(src);
    v[0] = f.x; v[1] = f.y; v[2] = f.z; v[3] = f.w;
  } else {
    v[0] = src[0]; v[1] = v[2] = v[3] = 0.f;
  }
}
template <bool VEC>
__device__ __forceinline__ void st_elems(float* __restrict__ dst, const float (&v)[4]) {
  if (VEC) st_f4(dst, make_float4(v[0], v[1], v[2], v[3]));
  else dst[0] = v[0];
}

struct Window {
  int64_t s;        // row id at position base + lane (-1 behind the end of the list)
  int32_t pm;       // perm[base + lane]
  unsigned bounds;  // lanes at which a new segment starts (or the list ends)
  unsigned todo;    // segment heads this warp processes: the heads among its first `stride` lanes
};

// A warp LOADS 32 positions and OWNS the first `stride` of them (consecutive warps overlap by 32 - stride positions; the
// index list is a few hundred kB and stays in L2).  With stride <= 16 every member of a segment of at most kLongSeg rows
// whose head the warp owns lies inside the loaded window.
__device__ __forceinline__ Window load_window(const int64_t* __restrict__ sorted, const int32_t* __restrict__ perm,
                                              int64_t R, int64_t base, int stride, int lane, bool skip_long) {
  Window w;
  const int64_t pos = base + lane;
  const bool valid = pos < R;
  w.s = valid ? sorted[pos] : -1;
  const int64_t s16 = pos + kLongSeg < R ? sorted[pos + kLongSeg] : -2;
  int64_t prev = (lane == 0 && base > 0) ? sorted[base - 1] : -1;
  w.pm = (valid && perm) ? perm[pos] : 0;
  const int64_t up = __shfl_up_sync(kFull, w.s, 1);
  if (lane > 0) prev = up;
  const bool head = valid && w.s != prev;
  const bool is_long = head && s16 == w.s;
  w.bounds = __ballot_sync(kFull, head || !valid);
  w.todo = __ballot_sync(kFull, head && lane < stride && !(skip_long && is_long));
  return w;
}

// acc += gradient rows of the window members q0 .. q1-1, in order, four loads in flight
template <bool VEC>
__device__ __forceinline__ void add_members(const GradSrc& gs, const Window& w, int q0, int q1, int col, bool active,
                                            float (&acc)[4]) {
  for (int q = q0; q < q1; q += 4) {
    float v[4][4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      v[u][0] = v[u][1] = v[u][2] = v[u][3] = 0.f;
      if (q + u < q1) {
        const int32_t pos = __shfl_sync(kFull, w.pm, q + u);
        if (active) ld_elems<VEC>(gs.row(pos) + col, v[u]);
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u)
      if (q + u < q1) {
#pragma unroll
        for (int e = 0; e < 4; ++e) acc[e] = __fadd_rn(acc[e], v[u][e]);
      }
  }
}

// acc += gradient rows of the positions j, j+1, ... while they belong to `row` (a segment that runs on behind the window
// and is not handled by the long role: only without a long-segment list)
template <bool VEC>
__device__ __forceinline__ void add_far(const GradSrc& gs, const int64_t* __restrict__ sorted, const int32_t* __restrict__ perm,
                                        int64_t j, int64_t R, int64_t row, int col, bool active, float (&acc)[4]) {
  while (j < R && sorted[j] == row) {
    float v[4];
    v[0] = v[1] = v[2] = v[3] = 0.f;
    if (active) ld_elems<VEC>(gs.row(perm[j]) + col, v);
#pragma unroll
    for (int e = 0; e < 4; ++e) acc[e] = __fadd_rn(acc[e], v[e]);
    ++j;
  }
}

// sum of the gradient rows of segment head lane h (first member already in acc)
template <bool VEC>
__device__ __forceinline__ void finish_segment(const GradSrc& gs, const Window& w, const int64_t* __restrict__ sorted,
                                               const int32_t* __restrict__ perm, int64_t base, int64_t R, int h, int64_t row,
                                               int col, bool active, float (&acc)[4]) {
  const unsigned above = w.bounds & ~((2u << h) - 1u);
  const int e = above ? __ffs(above) - 1 : 32;
  if (e > h + 1) add_members<VEC>(gs, w, h + 1, e, col, active, acc);
  if (e == 32) add_far<VEC>(gs, sorted, perm, base + 32, R, row, col, active, acc);   // runs on behind the window (rare)
}

// Block-cooperative sum of gradient rows [i, end) for one column group; the result is valid in warp 0.
template <bool VEC>
__device__ __forceinline__ void segment_sum_block(const GradSrc& gs, const int32_t* __restrict__ perm, int64_t i,
                                                  int64_t end, int col, bool active, float (*part)[128], float (&acc)[4]) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float a[4] = {0.f, 0.f, 0.f, 0.f};
  // this warp's members: j = i + warp + kRowWarps * k; lane l fetches the perm entry of member kb + l
  for (int64_t kb = 0; i + warp + kRowWarps * kb < end; kb += 32) {
    const int64_t jl = i + warp + kRowWarps * (kb + lane);
    const int32_t pl = jl < end ? perm[jl] : -1;
#pragma unroll
    for (int k0 = 0; k0 < 32; k0 += kLongUnroll) {
      if (i + warp + kRowWarps * (kb + k0) >= end) break;
      float v[kLongUnroll][4];
#pragma unroll
      for (int u = 0; u < kLongUnroll; ++u) {
        const int32_t pos = __shfl_sync(kFull, pl, k0 + u);
        v[u][0] = v[u][1] = v[u][2] = v[u][3] = 0.f;
        if (pos >= 0 && active) ld_elems<VEC>(gs.row(pos) + col, v[u]);
      }
#pragma unroll
      for (int u = 0; u < kLongUnroll; ++u)
#pragma unroll
        for (int e = 0; e < 4; ++e) a[e] = __fadd_rn(a[e], v[u][e]);
    }
  }
#pragma unroll
  for (int e = 0; e < 4; ++e) part[warp][lane * 4 + e] = a[e];
  __syncthreads();
  if (warp == 0) {
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      float t = 0.f;
#pragma unroll
      for (int w = 0; w < kRowWarps; ++w) t = __fadd_rn(t, part[w][lane * 4 + e]);
      acc[e] = t;
    }
  }
  __syncthreads();
}

template <bool VEC, int kHB>
__global__ void __launch_bounds__(kRowThreads, kHB == 2 ? 4 : 2) sparse_adam_rows_kernel(float* __restrict__ P, float* __restrict__ Mo,
                                                               float* __restrict__ Vo, int D,
                                                               const int64_t* __restrict__ sorted,
                                                               const int32_t* __restrict__ perm, int64_t R, GradSrc gs,
                                                               AdamScalars s, const float* __restrict__ scalars,
                                                               const ttam_step_state* __restrict__ st,
                                                               const int32_t* __restrict__ long_list, int long_blocks,
                                                               int stride) {
  __shared__ float part[kRowWarps][128];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (st) s.step_size = scalars[4 * st->step + 2];
  constexpr int W = VEC ? 128 : 32;
  constexpr int NE = VEC ? 4 : 1;
  if ((int)blockIdx.x < long_blocks) {  // ---- long role
    const int n = long_list[0];
    for (int e = blockIdx.x; e < n; e += long_blocks) {
      const int64_t i = long_list[1 + e];
      const int64_t row = sorted[i];
      const int64_t end = segment_end(sorted, i, R, row, lane);
      for (int c0 = 0; c0 < D; c0 += W) {
        const int col = c0 + lane * NE;
        const bool active = col < D;
        float g[4];
        segment_sum_block<VEC>(gs, perm, i, end, col, active, part, g);
        if (warp == 0 && active) {
          const int64_t off = row * (int64_t)D + col;
          float pp[4], mm[4], vv[4];
          ld_elems<VEC>(P + off, pp); ld_elems<VEC>(Mo + off, mm); ld_elems<VEC>(Vo + off, vv);
#pragma unroll
          for (int k = 0; k < NE; ++k) sparse_adam_elem(g[k], pp[k], mm[k], vv[k], s);
          st_elems<VEC>(P + off, pp); st_elems<VEC>(Mo + off, mm); st_elems<VEC>(Vo + off, vv);
        }
      }
    }
    return;
  }
  // ---- window role
  const int64_t base = (((int64_t)blockIdx.x - long_blocks) * kRowWarps + warp) * stride;
  if (base >= R) return;
  const Window w = load_window(sorted, perm, R, base, stride, lane, long_list != nullptr);
  for (int c0 = 0; c0 < D; c0 += W) {
    const int col = c0 + lane * NE;
    const bool active = col < D;
    unsigned todo = w.todo;
    while (todo) {
      int h[kHB];
      int64_t row[kHB];
      float pp[kHB][4], mm[kHB][4], vv[kHB][4], g[kHB][4];
#pragma unroll
      for (int u = 0; u < kHB; ++u) {
        h[u] = todo ? __ffs(todo) - 1 : -1;
        todo &= todo - 1;
        if (h[u] >= 0) {
          row[u] = __shfl_sync(kFull, w.s, h[u]);
          const int32_t pos = __shfl_sync(kFull, w.pm, h[u]);
          if (active) {
            const int64_t off = row[u] * (int64_t)D + col;
            ld_elems<VEC>(P + off, pp[u]); ld_elems<VEC>(Mo + off, mm[u]); ld_elems<VEC>(Vo + off, vv[u]);
            ld_elems<VEC>(gs.row(pos) + col, g[u]);
          }
        }
      }
#pragma unroll
      for (int u = 0; u < kHB; ++u) {
        if (h[u] < 0) break;
        finish_segment<VEC>(gs, w, sorted, perm, base, R, h[u], row[u], col, active, g[u]);
        if (active) {
          const int64_t off = row[u] * (int64_t)D + col;
#pragma unroll
          for (int k = 0; k < NE; ++k) sparse_adam_elem(g[u][k], pp[u][k], mm[u][k], vv[u][k], s);
          st_elems<VEC>(P + off, pp[u]); st_elems<VEC>(Mo + off, mm[u]); st_elems<VEC>(Vo + off, vv[u]);
        }
      }
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
// MUFU.RCP alone: div.approx (__fdividef) wraps it in a denormal-range test and two rescaling multiplies per element; the
// denominators here are >= eps
__device__ __forceinline__ float rcp_ftz(float x) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}

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
          pp[e] = fmaf(-a, __fmul_rn(mm[e], rcp_ftz(fmaf(w0[e], h, s.eps))), pp[e]);
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

// one lazily-updated row group held in registers: replay the zero-gradient steps (t_prev, step-1], then step `step`
template <int KIND, int NE>
__device__ __forceinline__ void lazy_apply(float (&pp)[4], float (&mm)[4], float (&vv)[4], const float (&g)[4], int t_prev,
                                           int step, const float* __restrict__ scalars, const AdamScalars& s) {
  const float step_size = KIND == TTAM_OPT_SGD ? 0.f : scalars[4 * step];
  const float bc2s = KIND == TTAM_OPT_SGD ? 1.f : scalars[4 * step + 1];
  float p1[NE], m1[NE], v1[NE];
#pragma unroll
  for (int e = 0; e < NE; ++e) { p1[e] = pp[e]; m1[e] = mm[e]; v1[e] = vv[e]; }
  replay<KIND, NE>(p1, m1, v1, t_prev, step - 1, scalars, s);
#pragma unroll
  for (int e = 0; e < NE; ++e) {
    dense_elem<KIND>(g[e], p1[e], m1[e], v1[e], s, step_size, bc2s);
    pp[e] = p1[e]; mm[e] = m1[e]; vv[e] = v1[e];
  }
}

template <bool VEC>
__device__ __forceinline__ void ld_state(const float* __restrict__ P, const float* __restrict__ Mo, const float* __restrict__ Vo,
                                         int64_t off, bool has_m, bool has_v, float (&pp)[4], float (&mm)[4], float (&vv)[4]) {
  ld_elems<VEC>(P + off, pp);
  if (has_m) ld_elems<VEC>(Mo + off, mm); else mm[0] = mm[1] = mm[2] = mm[3] = 0.f;
  if (has_v) ld_elems<VEC>(Vo + off, vv); else vv[0] = vv[1] = vv[2] = vv[3] = 0.f;
}
template <bool VEC>
__device__ __forceinline__ void st_state(float* __restrict__ P, float* __restrict__ Mo, float* __restrict__ Vo, int64_t off,
                                         bool has_m, bool has_v, const float (&pp)[4], const float (&mm)[4], const float (&vv)[4]) {
  st_elems<VEC>(P + off, pp);
  if (has_m) st_elems<VEC>(Mo + off, mm);
  if (has_v) st_elems<VEC>(Vo + off, vv);
}

template <int KIND, bool VEC, int kHB>
__global__ void __launch_bounds__(kRowThreads, kHB == 2 ? 3 : 2) lazy_rows_kernel(float* __restrict__ P, float* __restrict__ Mo,
                                                        float* __restrict__ Vo, int32_t* __restrict__ last_step, int D,
                                                        const int64_t* __restrict__ sorted,
                                                        const int32_t* __restrict__ perm, int64_t R, GradSrc gs,
                                                        const float* __restrict__ scalars, AdamScalars s, int step,
                                                        const ttam_step_state* __restrict__ st,
                                                        const int32_t* __restrict__ long_list, int long_blocks,
                                                        int stride) {
  __shared__ float part[kRowWarps][128];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (st) step = st->step;
  const bool has_v = KIND != TTAM_OPT_SGD;
  const bool has_m = KIND != TTAM_OPT_SGD || s.momentum != 0.f;
  constexpr int W = VEC ? 128 : 32;
  constexpr int NE = VEC ? 4 : 1;
  if ((int)blockIdx.x < long_blocks) {  // ---- long role
    const int n = long_list[0];
    for (int e = blockIdx.x; e < n; e += long_blocks) {
      const int64_t i = long_list[1 + e];
      const int64_t row = sorted[i];
      const int t_prev = last_step[row];
      const int64_t end = segment_end(sorted, i, R, row, lane);
      for (int c0 = 0; c0 < D; c0 += W) {
        const int col = c0 + lane * NE;
        const bool active = col < D;
        float g[4];
        segment_sum_block<VEC>(gs, perm, i, end, col, active, part, g);
        if (warp == 0 && active) {
          const int64_t off = row * (int64_t)D + col;
          float pp[4], mm[4], vv[4];
          ld_state<VEC>(P, Mo, Vo, off, has_m, has_v, pp, mm, vv);
          lazy_apply<KIND, NE>(pp, mm, vv, g, t_prev, step, scalars, s);
          st_state<VEC>(P, Mo, Vo, off, has_m, has_v, pp, mm, vv);
        }
      }
      __syncthreads();  // every warp has read last_step[row] before it changes
      if (threadIdx.x == 0) last_step[row] = step;
    }
    return;
  }
  // ---- window role
  const int64_t base = (((int64_t)blockIdx.x - long_blocks) * kRowWarps + warp) * stride;
  if (base >= R) return;
  const Window w = load_window(sorted, perm, R, base, stride, lane, long_list != nullptr);
  const bool mine = (w.todo >> lane) & 1u;
  const int tp_lane = mine ? last_step[w.s] : 0;
  for (int c0 = 0; c0 < D; c0 += W) {
    const int col = c0 + lane * NE;
    const bool active = col < D;
    unsigned todo = w.todo;
    while (todo) {
      int h[kHB];
      int64_t row[kHB];
      float pp[kHB][4], mm[kHB][4], vv[kHB][4], g[kHB][4];
#pragma unroll
      for (int u = 0; u < kHB; ++u) {
        h[u] = todo ? __ffs(todo) - 1 : -1;
        todo &= todo - 1;
        if (h[u] >= 0) {
          row[u] = __shfl_sync(kFull, w.s, h[u]);
          const int32_t pos = __shfl_sync(kFull, w.pm, h[u]);
          if (active) {
            ld_state<VEC>(P, Mo, Vo, row[u] * (int64_t)D + col, has_m, has_v, pp[u], mm[u], vv[u]);
            ld_elems<VEC>(gs.row(pos) + col, g[u]);
          }
        }
      }
#pragma unroll
      for (int u = 0; u < kHB; ++u) {
        if (h[u] < 0) break;
        finish_segment<VEC>(gs, w, sorted, perm, base, R, h[u], row[u], col, active, g[u]);
        const int t_prev = __shfl_sync(kFull, tp_lane, h[u]);
        if (active) {
          lazy_apply<KIND, NE>(pp[u], mm[u], vv[u], g[u], t_prev, step, scalars, s);
          st_state<VEC>(P, Mo, Vo, row[u] * (int64_t)D + col, has_m, has_v, pp[u], mm[u], vv[u]);
        }
      }
    }
  }
  if (mine) last_step[w.s] = step;
}

// Bring the unique rows of `sorted` up to step-1 (zero-gradient replay) so that the forward pass of step `step`
// reads current values; the row-wise update at the end of the step then finds nothing left to replay.
// The replay loop is arithmetic (a row that is touched every ~50 steps replays ~50 steps x D elements), so the row is
// spread over all 32 lanes: lane l holds the NE = D/32 elements l, l + 32, ... (coalesced 128-byte accesses) instead of
// four consecutive ones in 24 of the 32 lanes.
template <int KIND, int NE>
__global__ void __launch_bounds__(kRowThreads) lazy_catchup_kernel(float* __restrict__ P, float* __restrict__ Mo,
                                                           float* __restrict__ Vo, int32_t* __restrict__ last_step, int D,
                                                           const int64_t* __restrict__ sorted, int64_t R,
                                                           const float* __restrict__ scalars, AdamScalars s, int step,
                                                           const ttam_step_state* __restrict__ st) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t base = ((int64_t)blockIdx.x * kRowWarps + warp) * kCatchStride;
  if (base >= R) return;
  if (st) step = st->step;
  const bool has_v = KIND != TTAM_OPT_SGD;
  const bool has_m = KIND != TTAM_OPT_SGD || s.momentum != 0.f;
  constexpr int W = 32 * NE;
  // heads of the window (lane-parallel), then the ones that are behind
  const int64_t pos = base + lane;
  const bool valid = pos < R && lane < kCatchStride;
  const int64_t srow = valid ? sorted[pos] : -1;
  int64_t prev = (lane == 0 && base > 0) ? sorted[base - 1] : -1;
  const int64_t up = __shfl_up_sync(kFull, srow, 1);
  if (lane > 0) prev = up;
  const bool head = valid && srow != prev;
  const int tp_lane = head ? last_step[srow] : 0;
  const bool mine = head && tp_lane < step - 1;
  const unsigned all = __ballot_sync(kFull, mine);
  for (int c0 = 0; c0 < D; c0 += W) {
    const int col = c0 + lane;
    unsigned todo = all;
    while (todo) {
      int h[kCatchHB];
      int64_t off[kCatchHB];
      float pp[kCatchHB][NE], mm[kCatchHB][NE], vv[kCatchHB][NE];
#pragma unroll
      for (int u = 0; u < kCatchHB; ++u) {
        h[u] = todo ? __ffs(todo) - 1 : -1;
        todo &= todo - 1;
        if (h[u] >= 0) {
          off[u] = __shfl_sync(kFull, srow, h[u]) * (int64_t)D + col;
#pragma unroll
          for (int e = 0; e < NE; ++e) {
            const bool ok = col + 32 * e < D;
            pp[u][e] = ok ? P[off[u] + 32 * e] : 0.f;
            mm[u][e] = (ok && has_m) ? Mo[off[u] + 32 * e] : 0.f;
            vv[u][e] = (ok && has_v) ? Vo[off[u] + 32 * e] : 0.f;
          }
        }
      }
#pragma unroll
      for (int u = 0; u < kCatchHB; ++u) {
        if (h[u] < 0) break;
        const int t_prev = __shfl_sync(kFull, tp_lane, h[u]);
        replay<KIND, NE>(pp[u], mm[u], vv[u], t_prev, step - 1, scalars, s);
#pragma unroll
        for (int e = 0; e < NE; ++e) {
          if (col + 32 * e < D) {
            P[off[u] + 32 * e] = pp[u][e];
            if (has_m) Mo[off[u] + 32 * e] = mm[u][e];
            if (has_v) Vo[off[u] + 32 * e] = vv[u][e];
          }
        }
      }
    }
  }
  if (mine) last_step[srow] = step - 1;
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

// positions a warp owns (TTAM_ROWS_STRIDE: 8 | 16 | 32 for experiments; with a long-segment list 16 keeps every member of a
// segment inside the window a warp loads)
static int row_stride(bool with_list) {
  static const int env = [] { const char* e = getenv("TTAM_ROWS_STRIDE"); const int v = e ? atoi(e) : 0; return (v == 8 || v == 16 || v == 32) ? v : 0; }();
  return env ? env : (with_list ? 16 : 32);
}

static int rows_hb() {   // segment heads in flight per warp (TTAM_ROWS_HB = 2 | 4; 2 leaves room for twice the warps)
  static const int v = [] { const char* e = getenv("TTAM_ROWS_HB"); return (e && atoi(e) == 4) ? 4 : 2; }();
  return v;
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
  const bool vec = vec_ok(D, p, m, v, grad_a, ld_a, grad_b, ld_b);
  const int long_blocks = long_list ? num_sms() : 0;   // the long role of the same launch (most exit at once: few long segments)
  const int stride = row_stride(long_list != nullptr);
  const int blocks = long_blocks + (int)ceil_div(R, (int64_t)stride * kRowWarps);
  cudaStream_t cs = (cudaStream_t)stream;
#define SA_ARGS p, m, v, (int)D, sorted_idx, perm, R, gs, s, scalars, state_dev, long_list, long_blocks, stride
  if (!vec) sparse_adam_rows_kernel<false, 4><<<blocks, kRowThreads, 0, cs>>>(SA_ARGS);
  else if (rows_hb() == 2) sparse_adam_rows_kernel<true, 2><<<blocks, kRowThreads, 0, cs>>>(SA_ARGS);
  else sparse_adam_rows_kernel<true, 4><<<blocks, kRowThreads, 0, cs>>>(SA_ARGS);
#undef SA_ARGS
  TTAM_LAUNCH_CHECK();
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
  cudaStream_t st = (cudaStream_t)stream;
  const bool vec = vec_ok(D, p, m, v, grad_a, ld_a, grad_b, ld_b);
  const int long_blocks = long_list ? num_sms() : 0;
  const int stride = row_stride(long_list != nullptr);
  const int blocks = long_blocks + (int)ceil_div(R, (int64_t)stride * kRowWarps);
#define LZ_ARGS p, m, v, last_step, (int)D, sorted_idx, perm, R, gs, scalars, s, (int)step, state_dev, long_list, long_blocks, stride
#define CALL(K, V) lazy_rows_kernel<K, V, 4><<<blocks, kRowThreads, 0, st>>>(LZ_ARGS)
#define CALL2(K, V) lazy_rows_kernel<K, V, 2><<<blocks, kRowThreads, 0, st>>>(LZ_ARGS)
  if (!vec) TTAM_DISPATCH_KIND(kind, false, CALL);
  else if (rows_hb() == 2) TTAM_DISPATCH_KIND(kind, true, CALL2);
  else TTAM_DISPATCH_KIND(kind, true, CALL);
#undef CALL
#undef CALL2
#undef LZ_ARGS
  TTAM_LAUNCH_CHECK();
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
  const int blocks = (int)ceil_div(R, kCatchStride * kRowWarps);
  cudaStream_t st = (cudaStream_t)stream;
#define CALL(K, NE) lazy_catchup_kernel<K, NE><<<blocks, kRowThreads, 0, st>>>(p, m, v, last_step, (int)D, sorted_idx, R, scalars, s, (int)step, state_dev)
  // elements per lane: the whole row in one pass when D is 32, 64, 96 or 128; 128-column passes for wider rows
  const int ne = D % 128 == 0 ? 4 : (D % 32 == 0 && D < 128) ? (int)(D / 32) : D > 64 ? 4 : D > 32 ? 2 : 1;
  if (ne == 4) TTAM_DISPATCH_KIND(kind, 4, CALL);
  else if (ne == 3) TTAM_DISPATCH_KIND(kind, 3, CALL);
  else if (ne == 2) TTAM_DISPATCH_KIND(kind, 2, CALL);
  else TTAM_DISPATCH_KIND(kind, 1, CALL);
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
