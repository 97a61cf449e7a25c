// In-batch softmax loss (BASELINE.json north_star (2) / configs[1] "in-batch negatives"): an EXTENSION of the reference, whose
// only training loss is the sampled-negative BCE of training.py:770-803.  Definition = oracle/model.py
// inbatch_loss_forward_backward (pinned against torch autograd, "parity unpinned" with respect to the reference):
//     S = o_u o_p^T [B, B];   L_ce = mean_b (logsumexp_j S[b, j] - S[b, b]);   dL/dS = (softmax_rows(S) - I) / B
//     L = L_ce + lambda_u mean((q_u - sg t_p)^2) + lambda_i mean((q_p - sg t_u)^2)          (adaptive_mimic.py:66-67)
// One C call enqueues:
//   1. S        = o_u . o_p^T          ttam_linear_fwd  (tcgen05 TF32 tiles, or fp32 FFMA)      12.9 GFLOP at B = 8192, D = 96
//   2. rows     : lse_b, loss_b = lse_b - S_bb, and S <- dL/dS in place                       inbatch_rows_kernel (one pass + L2)
//   3. do_u     = dL/dS . o_p          ttam_linear_dgrad
//   4. do_p     = dL/dS^T . o_u        ttam_linear_wgrad (deterministic split over the batch)
//   5. mimic terms, dq = do + mimic gradient, the four loss scalars                          inbatch_finish_kernel(s)
// S lives in the caller's workspace (B^2 floats: 268 MB at B = 8192) and is read twice after step 2; a flash-style variant that
// keeps the score tiles in TMEM is the obvious next step, this one reuses the GEMM kernels of the towers.
#include "common.cuh"

namespace ttam {
namespace inbatch {

constexpr int kRowThreads = 256;

__device__ __forceinline__ float block_max(float v, float* sh) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
  __syncthreads();
  float r = sh[0];
  for (int w = 1; w < (int)(blockDim.x >> 5); ++w) r = fmaxf(r, sh[w]);
  __syncthreads();
  return r;
}
__device__ __forceinline__ float block_sum(float v, float* sh) {  // fixed shuffle tree + fixed warp order: deterministic
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
  __syncthreads();
  float r = 0.f;
  for (int w = 0; w < (int)(blockDim.x >> 5); ++w) r += sh[w];
  __syncthreads();
  return r;
}

// one block per row b of S
__global__ void __launch_bounds__(kRowThreads) inbatch_rows_kernel(float* __restrict__ S, int64_t ldS, int64_t B,
                                                                   float* __restrict__ rowloss, int backward) {
  __shared__ float sh[kRowThreads / 32];
  const int64_t b = blockIdx.x;
  float* row = S + b * ldS;   // ldS is a multiple of 4 floats: every row starts 16-byte aligned
  float m = -INFINITY;
  for (int64_t j = threadIdx.x * 4; j < B; j += kRowThreads * 4) {
    if (j + 3 < B) {
      const float4 v = ld_f4(row + j);
      m = fmaxf(m, fmaxf(fmaxf(v.x, v.y), fmaxf(v.z, v.w)));
    } else {
      for (int64_t e = j; e < B; ++e) m = fmaxf(m, row[e]);
    }
  }
  m = block_max(m, sh);
  float z = 0.f;
  for (int64_t j = threadIdx.x * 4; j < B; j += kRowThreads * 4) {
    if (j + 3 < B) {
      const float4 v = ld_f4(row + j);
      z += __expf(v.x - m) + __expf(v.y - m) + __expf(v.z - m) + __expf(v.w - m);
    } else {
      for (int64_t e = j; e < B; ++e) z += __expf(row[e] - m);
    }
  }
  z = block_sum(z, sh);
  const float lse = logf(z) + m;
  if (threadIdx.x == 0) rowloss[b] = lse - row[b];
  __syncthreads();  // row[b] is read before anyone rewrites it
  if (!backward) return;
  const float invB = 1.f / (float)B;
  for (int64_t j = threadIdx.x * 4; j < B; j += kRowThreads * 4) {
    if (j + 3 < B) {
      float4 v = ld_f4(row + j);
      v.x = __expf(v.x - lse); v.y = __expf(v.y - lse); v.z = __expf(v.z - lse); v.w = __expf(v.w - lse);
      if (b >= j && b < j + 4) {
        const int e = (int)(b - j);
        if (e == 0) v.x -= 1.f; else if (e == 1) v.y -= 1.f; else if (e == 2) v.z -= 1.f; else v.w -= 1.f;
      }
      v.x *= invB; v.y *= invB; v.z *= invB; v.w *= invB;
      st_f4(row + j, v);
    } else {
      for (int64_t e = j; e < B; ++e) row[e] = (__expf(row[e] - lse) - (e == b ? 1.f : 0.f)) * invB;
    }
  }
}

// dq_u = do_u + cu (q_u - t_p);  dq_p = do_p + ci (q_p - t_u);  per-block partial sums of the two squared errors
__global__ void __launch_bounds__(256) inbatch_mimic_kernel(const float* __restrict__ t_u, const float* __restrict__ t_p,
                                                            const float* __restrict__ q_u, const float* __restrict__ q_p,
                                                            const float* __restrict__ do_u, const float* __restrict__ do_p,
                                                            float* __restrict__ dq_u, float* __restrict__ dq_p, int64_t n, float cu,
                                                            float ci, float* __restrict__ partial) {
  __shared__ float sh[8];
  float su = 0.f, si = 0.f;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float du = q_u[i] - t_p[i], di = q_p[i] - t_u[i];
    su = fmaf(du, du, su);
    si = fmaf(di, di, si);
    if (dq_u) dq_u[i] = do_u[i] + cu * du;
    if (dq_p) dq_p[i] = do_p[i] + ci * di;
  }
  su = block_sum(su, sh);
  si = block_sum(si, sh);
  if (threadIdx.x == 0) {
    partial[2 * blockIdx.x] = su;
    partial[2 * blockIdx.x + 1] = si;
  }
}

// loss_out = {total, ce, mimic_user, mimic_item}
__global__ void __launch_bounds__(256) inbatch_finish_kernel(const float* __restrict__ rowloss, int64_t B, const float* __restrict__ partial,
                                                             int nblocks, int64_t BD, float lambda_u, float lambda_i, int mimic,
                                                             float* __restrict__ loss_out) {
  __shared__ float sh[8];
  float s = 0.f;
  for (int64_t i = threadIdx.x; i < B; i += blockDim.x) s += rowloss[i];
  s = block_sum(s, sh);
  float su = 0.f, si = 0.f;
  if (mimic) {
    for (int i = threadIdx.x; i < nblocks; i += blockDim.x) {
      su += partial[2 * i];
      si += partial[2 * i + 1];
    }
    su = block_sum(su, sh);
    si = block_sum(si, sh);
  }
  if (threadIdx.x == 0) {
    const float ce = s / (float)B;
    const float mu = su / (float)BD, mi = si / (float)BD;
    float total = ce;
    if (mimic && lambda_u > 0.f) total += lambda_u * mu;
    if (mimic && lambda_i > 0.f) total += lambda_i * mi;
    loss_out[0] = total;
    loss_out[1] = ce;
    loss_out[2] = mimic ? mu : 0.f;
    loss_out[3] = mimic ? mi : 0.f;
  }
}

constexpr int kMimicBlocks = 592;

}  // namespace inbatch
}  // namespace ttam

using namespace ttam;
using namespace ttam::inbatch;

extern "C" int64_t ttam_inbatch_loss_workspace_bytes(int64_t B, int64_t D) {
  if (B <= 0 || D <= 0) return 256;
  return align_up(B * align_up(B, 4) * 4, 256) + align_up(B * 4, 256) + align_up(kMimicBlocks * 2 * 4, 256) +
         align_up(ttam_linear_wgrad_workspace_bytes(B, B, D), 256) + 256;
}

extern "C" int ttam_inbatch_loss_fwd_bwd(const float* o_u, const float* o_p, const float* t_u, const float* t_p, const float* q_u,
                                         const float* q_p, float lambda_u, float lambda_i, float* loss_out, float* do_u, float* do_p,
                                         float* dq_u, float* dq_p, int64_t B, int64_t D, int precision, void* workspace,
                                         int64_t workspace_bytes, void* stream) {
  TTAM_CHECK_ARG(o_u && o_p && loss_out && workspace, "inbatch_loss: null pointer");
  TTAM_CHECK_ARG(B > 0 && D > 0 && B < (1ll << 31), "inbatch_loss: bad shape");
  const bool mimic = t_u && t_p && q_u && q_p;
  TTAM_CHECK_ARG(mimic || !(t_u || t_p || q_u || q_p), "inbatch_loss: pass all of t_u, t_p, q_u, q_p or none");
  const bool backward = do_u != nullptr;
  TTAM_CHECK_ARG(!backward || do_p, "inbatch_loss: do_u and do_p come together");
  TTAM_CHECK_ARG(!backward || !mimic || (dq_u && dq_p), "inbatch_loss: the mimic backward needs dq_u and dq_p");
  if (workspace_bytes < ttam_inbatch_loss_workspace_bytes(B, D)) {
    set_error("inbatch_loss: workspace too small");
    return TTAM_EWORKSPACE;
  }
  cudaStream_t s = (cudaStream_t)stream;
  char* w = (char*)workspace;
  const int64_t ldS = align_up(B, 4);
  float* S = (float*)w;                 w += align_up(B * ldS * 4, 256);
  float* rowloss = (float*)w;           w += align_up(B * 4, 256);
  float* partial = (float*)w;           w += align_up(kMimicBlocks * 2 * 4, 256);
  void* wg_ws = (void*)w;
  const int64_t wg_bytes = workspace_bytes - (w - (char*)workspace);
  // 1. S = o_u . o_p^T
  int rc = ttam_linear_fwd(o_u, D, nullptr, o_p, D, nullptr, S, ldS, B, B, D, TTAM_ACT_NONE, 0.f, 0, 0, nullptr, precision & 0xFF, stream);
  if (rc != TTAM_OK) return rc;
  // 2. row statistics, per-row loss, S <- dL/dS
  inbatch_rows_kernel<<<(unsigned)B, kRowThreads, 0, s>>>(S, ldS, B, rowloss, backward ? 1 : 0);
  TTAM_LAUNCH_CHECK();
  if (backward) {
    // 3. do_u = dL/dS . o_p          (dx[M,K] = dy[M,N] . w[N,K] with M = N = B, K = D)
    rc = ttam_linear_dgrad(S, ldS, o_p, do_u, D, nullptr, 0, 0, 1.f, 0, B, B, D, precision & 0xFF, stream);
    if (rc != TTAM_OK) return rc;
    // 4. do_p = dL/dS^T . o_u        (dw[N,K] = dy[M,N]^T . x[M,K])
    rc = ttam_linear_wgrad(S, ldS, o_u, D, nullptr, do_p, nullptr, B, B, D, 0, wg_ws, wg_bytes, precision & 0xFF, stream);
    if (rc != TTAM_OK) return rc;
  }
  // 5. mimic terms and the loss scalars
  int nblocks = 0;
  if (mimic) {
    const int64_t n = B * D;
    nblocks = (int)std::min<int64_t>(kMimicBlocks, ceil_div(n, 256));
    const float cu = lambda_u > 0.f ? lambda_u * 2.f / (float)n : 0.f;
    const float ci = lambda_i > 0.f ? lambda_i * 2.f / (float)n : 0.f;
    inbatch_mimic_kernel<<<nblocks, 256, 0, s>>>(t_u, t_p, q_u, q_p, do_u, do_p, backward ? dq_u : nullptr, backward ? dq_p : nullptr, n,
                                                 cu, ci, partial);
    TTAM_LAUNCH_CHECK();
  }
  inbatch_finish_kernel<<<1, 256, 0, s>>>(rowloss, B, partial, nblocks, B * D, lambda_u, lambda_i, mimic ? 1 : 0, loss_out);
  TTAM_LAUNCH_CHECK();
  return TTAM_OK;
}
