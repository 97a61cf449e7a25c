// fp32 SIMT GEMM core shared by gemm_simt.cu (linear layers) and topk.cu (exact scoring).
// EXACT = products rounded to fp32 and accumulated sequentially over k without contraction: the
// canonical score order of the retrieval path (oracle/retrieval.py canonical_scores).
#pragma once
#include "common.cuh"

namespace ttam {

constexpr int BM = 64, BN = 64, BK = 16, PAD = 4;

struct GemmP {
  const float* A;
  const float* B;
  float* C;
  int64_t lda, ldb, ldc;
  const int64_t* gatherA;  // row map for a K-contiguous A
  const int64_t* gatherK;  // k map for an MN-contiguous B (wgrad: rows of x)
  int M, N, K;
  int k_begin_stride;      // split-K chunk (0 = no split)
  const float* bias;
  int act;
  float dropout_p;
  uint64_t seed, offset;
  const ttam_step_state* st;  // device step state (dropout counter base), nullable
  const float* aux;
  int64_t ldaux;
  int mask_mode;
  float scale;
  int accumulate;
};

__device__ __forceinline__ float apply_act(int act, float x) {
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

// A_KC / B_KC: operand is K-contiguous (element (r,k) at base + r*ld + k) or
// MN-contiguous (element (r,k) at base + k*ld + r).
template <bool A_KC, bool B_KC, bool EXACT>
__global__ void __launch_bounds__(256) gemm_f32_kernel(GemmP p) {
  __shared__ __align__(16) float As[BK][BM + PAD];
  __shared__ __align__(16) float Bs[BK][BN + PAD];
  const int t = threadIdx.x;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  int k_lo = 0, k_hi = p.K;
  float* C = p.C;
  if (p.k_begin_stride > 0) {
    k_lo = blockIdx.z * p.k_begin_stride;
    k_hi = min(p.K, k_lo + p.k_begin_stride);
    C += (int64_t)blockIdx.z * p.M * p.ldc;  // partial buffer [split][M][ldc]
  }
  const int ty = t >> 4, tx = t & 15;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  // per-thread load coordinates
  // K-contiguous: row = t/4, kq = (t%4)*4 ; MN-contiguous: k = t/16, rq = (t%16)*4
  int64_t a_row_off = -1;
  if (A_KC) {
    int r = m0 + (t >> 2);
    if (r < p.M) {
      int64_t src = p.gatherA ? p.gatherA[r] : (int64_t)r;
      a_row_off = src * p.lda;
    }
  }
  int64_t b_row_off = -1;
  if (B_KC) {
    int r = n0 + (t >> 2);
    if (r < p.N) b_row_off = (int64_t)r * p.ldb;
  }

  for (int k0 = k_lo; k0 < k_hi; k0 += BK) {
    // ---- A tile
    if (A_KC) {
      const int kq = k0 + (t & 3) * 4;
      float v[4] = {0.f, 0.f, 0.f, 0.f};
      if (a_row_off >= 0) {
        const float* src = p.A + a_row_off + kq;
        if (kq + 3 < k_hi && ((reinterpret_cast<uintptr_t>(src) & 15) == 0)) {
          float4 f = ld_f4(src);
          v[0] = f.x; v[1] = f.y; v[2] = f.z; v[3] = f.w;
        } else {
#pragma unroll
          for (int j = 0; j < 4; ++j)
            if (kq + j < k_hi) v[j] = src[j];
        }
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) As[(t & 3) * 4 + j][t >> 2] = v[j];
    } else {
      const int k = k0 + (t >> 4);
      const int rq = m0 + (t & 15) * 4;
      float v[4] = {0.f, 0.f, 0.f, 0.f};
      if (k < k_hi) {
        const float* src = p.A + (int64_t)k * p.lda + rq;
        if (rq + 3 < p.M && ((reinterpret_cast<uintptr_t>(src) & 15) == 0)) {
          float4 f = ld_f4(src);
          v[0] = f.x; v[1] = f.y; v[2] = f.z; v[3] = f.w;
        } else {
#pragma unroll
          for (int j = 0; j < 4; ++j)
            if (rq + j < p.M) v[j] = src[j];
        }
      }
      st_f4(&As[t >> 4][(t & 15) * 4], make_float4(v[0], v[1], v[2], v[3]));
    }
    // ---- B tile
    if (B_KC) {
      const int kq = k0 + (t & 3) * 4;
      float v[4] = {0.f, 0.f, 0.f, 0.f};
      if (b_row_off >= 0) {
        const float* src = p.B + b_row_off + kq;
        if (kq + 3 < k_hi && ((reinterpret_cast<uintptr_t>(src) & 15) == 0)) {
          float4 f = ld_f4(src);
          v[0] = f.x; v[1] = f.y; v[2] = f.z; v[3] = f.w;
        } else {
#pragma unroll
          for (int j = 0; j < 4; ++j)
            if (kq + j < k_hi) v[j] = src[j];
        }
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) Bs[(t & 3) * 4 + j][t >> 2] = v[j];
    } else {
      const int k = k0 + (t >> 4);
      const int rq = n0 + (t & 15) * 4;
      float v[4] = {0.f, 0.f, 0.f, 0.f};
      if (k < k_hi) {
        const int64_t krow = p.gatherK ? p.gatherK[k] : (int64_t)k;
        const float* src = p.B + krow * p.ldb + rq;
        if (rq + 3 < p.N && ((reinterpret_cast<uintptr_t>(src) & 15) == 0)) {
          float4 f = ld_f4(src);
          v[0] = f.x; v[1] = f.y; v[2] = f.z; v[3] = f.w;
        } else {
#pragma unroll
          for (int j = 0; j < 4; ++j)
            if (rq + j < p.N) v[j] = src[j];
        }
      }
      st_f4(&Bs[t >> 4][(t & 15) * 4], make_float4(v[0], v[1], v[2], v[3]));
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      float4 a = ld_f4(&As[kk][ty * 4]);
      float4 b = ld_f4(&Bs[kk][tx * 4]);
      const float av[4] = {a.x, a.y, a.z, a.w};
      const float bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j)
          acc[i][j] = EXACT ? __fadd_rn(acc[i][j], __fmul_rn(av[i], bv[j])) : fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }

  // ---- epilogue
  const float keep_scale = p.dropout_p > 0.f ? 1.f / (1.f - p.dropout_p) : 1.f;
  const uint64_t rng_base = p.offset + ((p.dropout_p > 0.f && p.st) ? p.st->rng_offset : 0ull);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + ty * 4 + i;
    if (m >= p.M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n >= p.N) continue;
      float v = acc[i][j];
      if (p.bias) v += p.bias[n];
      v = apply_act(p.act, v);
      if (p.dropout_p > 0.f) {
        v = dropout_keep(p.seed, rng_base + (uint64_t)m * (uint64_t)p.N + (uint64_t)n, p.dropout_p) ? v * keep_scale : 0.f;
      }
      if (p.mask_mode == 1) v = (p.aux[(int64_t)m * p.ldaux + n] > 0.f) ? v : 0.f;
      v *= p.scale;
      float* dst = C + (int64_t)m * p.ldc + n;
      if (p.accumulate) v += *dst;
      *dst = v;
    }
  }
}

}  // namespace ttam
