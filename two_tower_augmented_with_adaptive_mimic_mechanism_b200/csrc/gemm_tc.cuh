// TF32 tcgen05 implementations behind ttam_linear_fwd / _dgrad / _wgrad (TTAM_PREC_TF32); see gemm_tc.cu.
#pragma once
#include "gemm_simt.cuh"

namespace ttam {

// `prerounded`: bit 0 = x, bit 1 = w already hold TF32-representable values (TTAM_PREC_TF32_X / _XW): the kernel skips
// its in-place rounding pass for that operand.

int tc_linear_fwd(const float* x, int64_t ldx, const int64_t* gather, const float* w, int64_t ldw, const float* bias,
                  float* y, int64_t ldy, int64_t M, int64_t N, int64_t K, int act, float dropout_p, uint64_t seed,
                  uint64_t offset, const ttam_step_state* state_dev, int prerounded, cudaStream_t st);
int tc_linear_dgrad(const float* dy, int64_t lddy, const float* w, float* dx, int64_t lddx, const float* aux, int64_t ldaux,
                    int mask_mode, float scale, int accumulate, int64_t M, int64_t N, int64_t K, cudaStream_t st);
int tc_wgrad_splits(int64_t M, int64_t N, int64_t K);
// part[z][N][K] and (when colsum_partial is non-null) colsum_partial[z][N] = column sums of dy, for z < *real_splits
int tc_linear_wgrad_partials(const float* dy, int64_t lddy, const float* x, int64_t ldx, const int64_t* gather, float* partial,
                             float* colsum_partial, int64_t M, int64_t N, int64_t K, int* real_splits, int prerounded,
                             cudaStream_t st);

// gemm_tma.cu: C[M,N] = A[M,K] . B[N,K]^T with TMA-fed operands (B pre-rounded to TF32; A rounded in the kernel when roundA).
// Returns TTAM_OK, a negative error, or +1 when the operands do not qualify (alignment): the caller falls back to tc_linear_*.
int tma_gemm(const float* A, int64_t lda, const float* B, int64_t ldb, float* C, int64_t ldc, int64_t M, int64_t N, int64_t K,
             const float* bias, int relu, float dropout_p, uint64_t seed, uint64_t offset, const ttam_step_state* st,
             const float* aux, int64_t ldaux, int mask_mode, float scale, int accumulate, int roundA, int round_out,
             cudaStream_t stream);

// gemm_tma.cu: the weight gradient with TMA-fed MN-major operands (no row gather).  Same partial layout as
// tc_linear_wgrad_partials; +1 when the operands do not qualify.
int tma_wgrad_splits(int64_t M, int64_t N, int64_t K);
int tma_wgrad_partials(const float* dy, int64_t lddy, const float* x, int64_t ldx, float* partial, float* colsum_partial, int64_t M,
                       int64_t N, int64_t K, int* real_splits, int prerounded, cudaStream_t stream);


// gemm_tma.cu: the weight gradient of the bag-form layer 1 on the tensor cores (CSR rows expanded into the A stage).  Same
// partial layout; +1 when the shape is not covered.
int tma_bag_wgrad_splits(int64_t R, int64_t H, int64_t F);
int tma_bag_wgrad_partials(const int64_t* rowptr, const void* entries, const float* tail, int64_t T, int64_t tail_start,
                           const int64_t* gather, int64_t R, const float* dh, int64_t lddh, float* partial, float* colsum_partial,
                           int64_t H, int64_t F, int* real_splits, cudaStream_t stream);

}  // namespace ttam
