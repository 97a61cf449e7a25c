// fp32 SIMT GEMMs with fused gather / bias / activation / dropout / mask epilogues (TTAM_PREC_FP32).
// These keep the reference's fp32 arithmetic (encoders.py:121-144,157-162 -> ATen addmm); the tcgen05
// versions in gemm_tc.cu trade that for tensor-core throughput.
//
//   fwd   (NT): y[M,N]  = act(x[g(m),:K] . w[n,:K] + b[n])
//   dgrad (NN): dx[M,K] = (dy[M,:N] . w[:N,k]) * mask * scale   (+ dx)
//   wgrad (TN): dw[N,K] = sum_m dy[m,n] * x[g(m),k]              (split over m, deterministic reduce)
#include "gemm_simt.cuh"
#include "gemm_tc.cuh"

#define TTAM_TRY_RC(expr)             \
  do {                                \
    const int rc__ = (expr);          \
    if (rc__ != TTAM_OK) return rc__; \
  } while (0)

namespace ttam {

// out[i] (+)= sum_s partial[s][i] for two tensors at once (a weight gradient and its bias gradient), fixed order:
// eight lanes share one output element, lane l sums the splits l, l+8, ... (eight loads in flight per element instead
// of one serial chain per thread: the old one-thread-per-element loop was latency-bound at 13 us per launch), and the
// eight partial sums are combined by a fixed shuffle tree, so the result does not depend on the launch geometry.
__global__ void __launch_bounds__(256) splitk_reduce_kernel(const float* __restrict__ partial_a, int64_t numel_a,
                                                            float* __restrict__ out_a, const float* __restrict__ partial_b,
                                                            int64_t numel_b, float* __restrict__ out_b, int splits,
                                                            int accumulate) {
  const int64_t total = numel_a + numel_b;
  const int l = threadIdx.x & 7;
  for (int64_t e = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 3; e < ((total + 31) & ~31ll);
       e += ((int64_t)gridDim.x * blockDim.x) >> 3) {
    const bool live = e < total;
    const bool in_a = e < numel_a;
    const float* src = in_a ? partial_a + e : partial_b + (e - numel_a);
    const int64_t stride = in_a ? numel_a : numel_b;
    float s0 = 0.f, s1 = 0.f;
    if (live) {
      int k = l;
      for (; k + 8 < splits; k += 16) {
        s0 += src[(int64_t)k * stride];
        s1 += src[(int64_t)(k + 8) * stride];
      }
      if (k < splits) s0 += src[(int64_t)k * stride];
    }
    float s = s0 + s1;
    s += __shfl_xor_sync(0xffffffffu, s, 1);
    s += __shfl_xor_sync(0xffffffffu, s, 2);
    s += __shfl_xor_sync(0xffffffffu, s, 4);
    if (live && l == 0) {
      float* dst = in_a ? out_a + e : out_b + (e - numel_a);
      *dst = accumulate ? *dst + s : s;
    }
  }
}

static inline int reduce_blocks(int64_t total) {
  return (int)std::min<int64_t>(ceil_div(total * 8, 256), (int64_t)num_sms() * 16);
}

// Vector form (every count a multiple of 4, 16-byte aligned buffers): L lanes (a power of two, <= 32) share one group of FOUR
// consecutive output elements, lane l sums the splits l, l + L, ... with two independent chains, the L partial sums are
// combined by a fixed shuffle tree.  A warp load then reads (32 / L) * 16 contiguous bytes of each of L splits - whole
// sectors - where the scalar form above reads 16 bytes of each of 8 splits (half sectors): the 15 MB of partials of the
// [192, 608] layer-1 gradient took 47 us with it.  L is chosen by the host from the problem size only, so the summation
// order (and the result) does not depend on the launch geometry.
__global__ void __launch_bounds__(256) splitk_reduce_vec_kernel(const float* __restrict__ partial_a, int64_t numel_a,
                                                                float* __restrict__ out_a, const float* __restrict__ partial_b,
                                                                int64_t numel_b, float* __restrict__ out_b, int splits,
                                                                int accumulate, int L) {
  const int64_t groups = (numel_a + numel_b) >> 2;
  const int l = threadIdx.x & (L - 1);
  const int64_t gstride = ((int64_t)gridDim.x * blockDim.x) / L;
  const int64_t gmax = (groups + (32 / L) - 1) / (32 / L) * (32 / L);   // whole warps stay in the loop (shuffles)
  for (int64_t q = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / L; q < gmax; q += gstride) {
    const bool live = q < groups;
    const int64_t e = q << 2;
    const bool in_a = e < numel_a;
    const float* src = in_a ? partial_a + e : partial_b + (e - numel_a);
    const int64_t stride = in_a ? numel_a : numel_b;
    float4 s0 = make_float4(0.f, 0.f, 0.f, 0.f), s1 = s0;
    if (live) {
      int k = l;
      for (; k + L < splits; k += 2 * L) {
        const float4 a = ld_f4(src + (int64_t)k * stride), b = ld_f4(src + (int64_t)(k + L) * stride);
        s0.x += a.x; s0.y += a.y; s0.z += a.z; s0.w += a.w;
        s1.x += b.x; s1.y += b.y; s1.z += b.z; s1.w += b.w;
      }
      if (k < splits) {
        const float4 a = ld_f4(src + (int64_t)k * stride);
        s0.x += a.x; s0.y += a.y; s0.z += a.z; s0.w += a.w;
      }
    }
    float4 t = make_float4(s0.x + s1.x, s0.y + s1.y, s0.z + s1.z, s0.w + s1.w);
    for (int o = 1; o < L; o <<= 1) {
      t.x += __shfl_xor_sync(0xffffffffu, t.x, o); t.y += __shfl_xor_sync(0xffffffffu, t.y, o);
      t.z += __shfl_xor_sync(0xffffffffu, t.z, o); t.w += __shfl_xor_sync(0xffffffffu, t.w, o);
    }
    if (live && l == 0) {
      float* dst = in_a ? out_a + e : out_b + (e - numel_a);
      if (accumulate) {
        const float4 old = ld_f4(dst);
        t.x += old.x; t.y += old.y; t.z += old.z; t.w += old.w;
      }
      st_f4(dst, t);
    }
  }
}

// one entry point for both forms
static int launch_splitk_reduce(const float* pa, int64_t na, float* oa, const float* pb, int64_t nb, float* ob, int splits,
                                int accumulate, cudaStream_t s) {
  const bool vec = na % 4 == 0 && nb % 4 == 0 && (((uintptr_t)pa | (uintptr_t)oa | (uintptr_t)pb | (uintptr_t)ob) & 15) == 0;
  if (!vec) {
    splitk_reduce_kernel<<<reduce_blocks(na + nb), 256, 0, s>>>(pa, na, oa, pb, nb, ob, splits, accumulate);
  } else {
    const int64_t groups = (na + nb) / 4;
    int L = 1;
    while (L < 32 && 2 * L <= splits && groups * L < 75000) L <<= 1;   // enough threads to hide the load latency
    const int64_t threads = align_up(groups, 32 / L) * L;
    const int blocks = (int)std::min<int64_t>(ceil_div(threads, 256), (int64_t)num_sms() * 16);
    splitk_reduce_vec_kernel<<<blocks, 256, 0, s>>>(pa, na, oa, pb, nb, ob, splits, accumulate, L);
  }
  TTAM_LAUNCH_CHECK();
  return TTAM_OK;
}

// partial[s][n] = sum_{m in chunk s} dy[m][n].  One block per row chunk; a warp reads 32 consecutive columns of a row
// (128 B), the 8 warps take rows lo+0..7, lo+8..15, ...; the 8 partial sums are combined in warp order.
__global__ void __launch_bounds__(256) colsum_partial_kernel(const float* __restrict__ dy, int64_t lddy, int M, int N,
                                                             int chunk, float* __restrict__ partial) {
  __shared__ float red[8][33];
  const int s = blockIdx.x;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int lo = s * chunk, hi = min(M, lo + chunk);
  for (int n0 = 0; n0 < N; n0 += 32) {
    const int n = n0 + tx;
    float a0 = 0.f, a1 = 0.f;
    if (n < N) {
      int m = lo + ty;
      for (; m + 8 < hi; m += 16) {
        a0 += dy[(int64_t)m * lddy + n];
        a1 += dy[(int64_t)(m + 8) * lddy + n];
      }
      if (m < hi) a0 += dy[(int64_t)m * lddy + n];
    }
    red[ty][tx] = a0 + a1;
    __syncthreads();
    if (ty == 0 && n < N) {
      float t = 0.f;
#pragma unroll
      for (int w = 0; w < 8; ++w) t += red[w][tx];
      partial[(int64_t)s * N + n] = t;
    }
    __syncthreads();
  }
}

// out[n] (+)= sum_s partial[s][n]: one warp per column, lanes stride over the splits, fixed-order tree at the end
__global__ void __launch_bounds__(256) colsum_final_kernel(const float* __restrict__ partial, int splits, int N,
                                                           float* __restrict__ out, int accumulate) {
  const int n = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (n >= N) return;
  float a = 0.f;
  for (int s = lane; s < splits; s += 32) a += partial[(int64_t)s * N + n];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) a += __shfl_down_sync(0xffffffffu, a, o);
  if (lane == 0) out[n] = accumulate ? out[n] + a : a;
}

static int colsum_splits(int64_t M) {
  int64_t s = ceil_div(M, 32);
  const int64_t cap = (int64_t)num_sms() * 8;
  if (s > cap) s = cap;
  return (int)(s < 1 ? 1 : s);
}

static int wgrad_splits(int64_t M, int64_t N, int64_t K) {
  int64_t tiles = ceil_div(N, BM) * ceil_div(K, BN);
  int64_t want = ceil_div((int64_t)num_sms() * 4, tiles);
  int64_t max_by_rows = ceil_div(M, 256);
  int64_t s = want < max_by_rows ? want : max_by_rows;
  if (s < 1) s = 1;
  if (s > 1024) s = 1024;
  return (int)s;
}

}  // namespace ttam

using namespace ttam;

extern "C" int ttam_linear_fwd(const float* x, int64_t ldx, const int64_t* gather, const float* w, int64_t ldw, const float* bias,
                               float* y, int64_t ldy, int64_t M, int64_t N, int64_t K, int act, float dropout_p,
                               uint64_t seed, uint64_t offset, const ttam_step_state* state_dev, int precision,
                               void* stream) {
  TTAM_CHECK_ARG(x && w && y, "linear_fwd: null pointer");
  TTAM_CHECK_ARG(M >= 0 && N > 0 && K > 0 && ldx >= K && ldy >= N, "linear_fwd: bad shape");
  TTAM_CHECK_ARG(M < (1ll << 31) && N < (1ll << 31) && K < (1ll << 31), "linear_fwd: dimension too large");
  TTAM_CHECK_ARG(dropout_p >= 0.f && dropout_p < 1.f, "linear_fwd: dropout must be in [0,1)");
  TTAM_CHECK_ARG(act >= TTAM_ACT_NONE && act <= TTAM_ACT_SELU, "linear_fwd: unknown activation %d", act);
  if (M == 0) return TTAM_OK;
  TTAM_CHECK_ARG(ldw >= K, "linear_fwd: ldw < K");
  const int prerounded = (precision >> 8) & 3;  // TTAM_PREC_X_ROUNDED / TTAM_PREC_W_ROUNDED
  const int round_out = (precision & TTAM_PREC_OUT_ROUNDED) ? 1 : 0;
  precision &= 0xFF;
  if (precision == TTAM_PREC_TF32) {
    if (!gather && (prerounded & 2) && (act == TTAM_ACT_NONE || act == TTAM_ACT_RELU)) {
      // plain K-major operands with a pre-rounded weight: the TMA-fed persistent kernel (gemm_tma.cu)
      const int rc = tma_gemm(x, ldx, w, ldw, y, ldy, M, N, K, bias, act == TTAM_ACT_RELU, dropout_p, seed, offset, state_dev,
                              nullptr, 0, 0, 1.f, 0, !(prerounded & 1), round_out, (cudaStream_t)stream);
      if (rc <= 0) return rc;
    }
    const int rc = tc_linear_fwd(x, ldx, gather, w, ldw, bias, y, ldy, M, N, K, act, dropout_p, seed, offset, state_dev, prerounded,
                                 (cudaStream_t)stream);
    if (rc != TTAM_OK || !round_out) return rc;
    return ttam_round_tf32(y, ldy, M, N, stream);
  }
  if (precision != TTAM_PREC_FP32) {
    set_error("linear_fwd: precision %d is not built into this library", precision);
    return TTAM_EUNSUPPORTED;
  }
  GemmP p{};
  p.A = x; p.B = w; p.C = y; p.lda = ldx; p.ldb = ldw; p.ldc = ldy; p.gatherA = gather;
  p.M = (int)M; p.N = (int)N; p.K = (int)K; p.bias = bias; p.act = act; p.dropout_p = dropout_p;
  p.seed = seed; p.offset = offset; p.st = state_dev; p.scale = 1.f;
  dim3 grid((unsigned)ceil_div(N, BN), (unsigned)ceil_div(M, BM), 1);
  gemm_f32_kernel<true, true, false><<<grid, 256, 0, (cudaStream_t)stream>>>(p);
  TTAM_LAUNCH_CHECK();
  return TTAM_OK;
}

extern "C" int ttam_linear_dgrad(const float* dy, int64_t lddy, const float* w, float* dx, int64_t lddx,
                                 const float* aux, int64_t ldaux, int mask_mode, float scale, int accumulate,
                                 int64_t M, int64_t N, int64_t K, int precision, void* stream) {
  TTAM_CHECK_ARG(dy && w && dx, "linear_dgrad: null pointer");
  TTAM_CHECK_ARG(M >= 0 && N > 0 && K > 0 && lddy >= N && lddx >= K, "linear_dgrad: bad shape");
  TTAM_CHECK_ARG(mask_mode == 0 || (mask_mode == 1 && aux && ldaux >= K), "linear_dgrad: bad mask arguments");
  if (M == 0) return TTAM_OK;
  const int flags = precision;
  precision &= 0xFF;
  if (flags & TTAM_PREC_WT) {
    // w is the TF32-rounded TRANSPOSED weight [K, N] (ttam_prepare_weights): dx = dy . w is then the K-major x K-major
    // product of the TMA-fed kernel
    TTAM_CHECK_ARG(precision == TTAM_PREC_TF32, "linear_dgrad: TTAM_PREC_WT needs TTAM_PREC_TF32");
    const int rc = tma_gemm(dy, lddy, w, N, dx, lddx, M, K, N, nullptr, 0, 0.f, 0, 0, nullptr, aux, ldaux, mask_mode, scale, accumulate,
                            !(flags & TTAM_PREC_X_ROUNDED), (flags & TTAM_PREC_OUT_ROUNDED) ? 1 : 0, (cudaStream_t)stream);
    if (rc == 1) {
      set_error("linear_dgrad: TTAM_PREC_WT needs 16-byte aligned operands and N %% 4 == 0");
      return TTAM_EINVAL;
    }
    return rc;
  }
  if (precision == TTAM_PREC_TF32) {
    const int rc = tc_linear_dgrad(dy, lddy, w, dx, lddx, aux, ldaux, mask_mode, scale, accumulate, M, N, K, (cudaStream_t)stream);
    if (rc != TTAM_OK || !(flags & TTAM_PREC_OUT_ROUNDED)) return rc;
    return ttam_round_tf32(dx, lddx, M, K, stream);
  }
  if (precision != TTAM_PREC_FP32) {
    set_error("linear_dgrad: precision %d is not built into this library", precision);
    return TTAM_EUNSUPPORTED;
  }
  // C[M,K] = A[M,N] . B where B(row=k, kk=n) = w[n*K + k]  (MN-contiguous)
  GemmP p{};
  p.A = dy; p.B = w; p.C = dx; p.lda = lddy; p.ldb = K; p.ldc = lddx;
  p.M = (int)M; p.N = (int)K; p.K = (int)N; p.aux = aux; p.ldaux = ldaux; p.mask_mode = mask_mode;
  p.scale = scale; p.accumulate = accumulate;
  dim3 grid((unsigned)ceil_div(K, BN), (unsigned)ceil_div(M, BM), 1);
  gemm_f32_kernel<true, false, false><<<grid, 256, 0, (cudaStream_t)stream>>>(p);
  TTAM_LAUNCH_CHECK();
  return TTAM_OK;
}

extern "C" int64_t ttam_linear_wgrad_workspace_bytes(int64_t M, int64_t N, int64_t K) {
  int s = wgrad_splits(M, N, K);  // the workspace serves either precision / kernel
  const int s_tc = tc_wgrad_splits(M, N, K), s_tma = tma_wgrad_splits(M, N, K);
  if (s_tc > s) s = s_tc;
  if (s_tma > s) s = s_tma;
  return ((int64_t)s * N * K + (int64_t)colsum_splits(M) * N) * (int64_t)sizeof(float);
}

extern "C" int ttam_linear_wgrad(const float* dy, int64_t lddy, const float* x, int64_t ldx, const int64_t* gather,
                                 float* dw, float* db, int64_t M, int64_t N, int64_t K, int accumulate,
                                 void* workspace, int64_t workspace_bytes, int precision, void* stream) {
  TTAM_CHECK_ARG(dy && x && dw && workspace, "linear_wgrad: null pointer");
  TTAM_CHECK_ARG(M > 0 && N > 0 && K > 0 && lddy >= N && ldx >= K, "linear_wgrad: bad shape");
  const int prerounded = (precision >> 8) & 3;
  precision &= 0xFF;
  if (precision != TTAM_PREC_FP32 && precision != TTAM_PREC_TF32) {
    set_error("linear_wgrad: precision %d is not built into this library", precision);
    return TTAM_EUNSUPPORTED;
  }
  if (workspace_bytes < ttam_linear_wgrad_workspace_bytes(M, N, K)) {
    set_error("linear_wgrad: workspace too small");
    return TTAM_EWORKSPACE;
  }
  cudaStream_t s = (cudaStream_t)stream;
  int splits = precision == TTAM_PREC_TF32 ? tc_wgrad_splits(M, N, K) : wgrad_splits(M, N, K);
  if (precision == TTAM_PREC_TF32 && tma_wgrad_splits(M, N, K) > splits) splits = tma_wgrad_splits(M, N, K);
  const int chunk = (int)align_up(ceil_div(M, splits), BK);
  float* partial_w = (float*)workspace;
  float* partial_b = partial_w + (int64_t)splits * N * K;
  if (precision == TTAM_PREC_TF32) {
    int real = 0;
    int rc = gather ? 1 : tma_wgrad_partials(dy, lddy, x, ldx, partial_w, db ? partial_b : nullptr, M, N, K, &real, prerounded, s);
    if (rc < 0) return rc;
    if (rc == 1) rc = tc_linear_wgrad_partials(dy, lddy, x, ldx, gather, partial_w, db ? partial_b : nullptr, M, N, K, &real, prerounded, s);
    if (rc != TTAM_OK) return rc;
    // one launch sums the weight-gradient partials and (the bias gradient came out of the same pass over dy) the
    // column-sum partials
    const int64_t numel = N * K, numel_b = db ? N : 0;
    return launch_splitk_reduce(partial_w, numel, dw, partial_b, numel_b, db, real, accumulate, s);
  } else {
  // C[N,K] = sum_m A(row=n, k=m) * B(row=k, k=m);  A = dy (MN-contiguous), B = x rows (MN-contiguous, gathered on m)
  GemmP p{};
  p.A = dy; p.B = x; p.C = partial_w; p.lda = lddy; p.ldb = ldx; p.ldc = K; p.gatherK = gather;
  p.M = (int)N; p.N = (int)K; p.K = (int)M; p.k_begin_stride = chunk; p.scale = 1.f;
  const int real_splits = (int)ceil_div(M, chunk);
  dim3 grid((unsigned)ceil_div(K, BN), (unsigned)ceil_div(N, BM), (unsigned)real_splits);
  gemm_f32_kernel<false, false, false><<<grid, 256, 0, s>>>(p);
  TTAM_LAUNCH_CHECK();
  {
    int64_t numel = N * K;
    TTAM_TRY_RC(launch_splitk_reduce(partial_w, numel, dw, nullptr, 0, nullptr, real_splits, accumulate, s));
  }
  }
  if (db) {
    const int cs = colsum_splits(M);
    const int cchunk = (int)ceil_div(M, cs);
    const int real_cs = (int)ceil_div(M, cchunk);
    colsum_partial_kernel<<<real_cs, 256, 0, s>>>(dy, lddy, (int)M, (int)N, cchunk, partial_b);
    TTAM_LAUNCH_CHECK();
    colsum_final_kernel<<<(int)ceil_div(N, 8), 256, 0, s>>>(partial_b, real_cs, (int)N, db, accumulate);
    TTAM_LAUNCH_CHECK();
  }
  return TTAM_OK;
}

// ---- bag-form layer 1, weight gradient on the tensor cores (csrc/gemm_tma.cu bag_wgrad_tc_kernel) ---------------------------
extern "C" int64_t ttam_bag_linear_wgrad_tc_workspace_bytes(int64_t R, int64_t H, int64_t F) {
  const int64_t s = tma_bag_wgrad_splits(R > 0 ? R : 1, H, F);
  return align_up((s * H * F + s * H) * (int64_t)sizeof(float), 256) + 256;
}

// Returns TTAM_OK, an error, or +1 when the shape is not covered (the caller then uses ttam_bag_linear_wgrad).
extern "C" int ttam_bag_linear_wgrad_tc(const int64_t* rowptr, const void* entries, const float* tail, int64_t T, int64_t tail_start,
                                        const int64_t* gather, int64_t R, const float* dh, int64_t lddh, float* dw, int64_t lddw,
                                        float* db, int64_t H, int64_t F, int accumulate, void* workspace, int64_t workspace_bytes,
                                        void* stream) {
  TTAM_CHECK_ARG(rowptr && entries && dh && dw && workspace, "bag_linear_wgrad_tc: null pointer");
  TTAM_CHECK_ARG(T >= 0 && T <= 8 && tail_start + T == F && (T == 0 || tail), "bag_linear_wgrad_tc: bad dense tail");
  if (R <= 0 || lddw != F) return 1;
  if (workspace_bytes < ttam_bag_linear_wgrad_tc_workspace_bytes(R, H, F)) {
    set_error("bag_linear_wgrad_tc: workspace too small");
    return TTAM_EWORKSPACE;
  }
  cudaStream_t s = (cudaStream_t)stream;
  const int splits = tma_bag_wgrad_splits(R, H, F);
  float* partial_w = (float*)workspace;
  float* partial_b = partial_w + (int64_t)splits * H * F;
  int real = 0;
  const int rc = tma_bag_wgrad_partials(rowptr, entries, tail, T, tail_start, gather, R, dh, lddh, partial_w, db ? partial_b : nullptr, H, F,
                                        &real, s);
  if (rc != TTAM_OK) return rc;
  return launch_splitk_reduce(partial_w, H * F, dw, partial_b, db ? H : 0, db, real, accumulate, s);
}
