/* ttam.h — C ABI of libttam.so: the B200 (sm_100a) hot path of the two-tower recommender.
 *
 * The reference (alperkartkaya2-afk/two-tower-augmented-with-adaptive-mimic-mechanism) is pure
 * Python/PyTorch and has no FFI layer, so there is no existing binding to mirror; each entry point
 * below replaces one chain of ATen calls on the reference's hot path and cites it (file:line relative
 * to the reference root).  INTEGRATION.md shows the ctypes stub a maintainer would add.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless the name ends in _host; row-major, contiguous rows
 *   - indices are int64 (reference: datasets.py:34-39, adaptive_mimic.py:101-102); floats are fp32
 *     unless the name says bf16 (uint16_t storage)
 *   - no allocation, no host synchronisation inside; work is enqueued on `stream` (a cudaStream_t
 *     passed as void*); callers provide outputs and workspaces (ttam_*_workspace_bytes tells how much)
 *   - return 0 on success, a negative TTAM_E* code otherwise; ttam_last_error() returns a
 *     thread-local message for the last failing call
 */
#ifndef TTAM_H_
#define TTAM_H_
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define TTAM_OK 0
#define TTAM_EINVAL -1      /* bad argument (shape, alignment, unsupported option) */
#define TTAM_ECUDA -2       /* a CUDA runtime call failed (launch error is sticky: see message) */
#define TTAM_EWORKSPACE -3  /* workspace too small */
#define TTAM_EUNSUPPORTED -4

/* activations of build_feature_encoder (encoders.py:68-78) */
#define TTAM_ACT_NONE 0
#define TTAM_ACT_RELU 1
#define TTAM_ACT_GELU 2
#define TTAM_ACT_TANH 3
#define TTAM_ACT_SELU 4

/* arithmetic of the GEMM-shaped ops */
#define TTAM_PREC_FP32 0 /* SIMT FFMA, fp32 in / fp32 accumulate (bit-faithful to the fp32 reference up to summation order) */
#define TTAM_PREC_TF32 1 /* tcgen05 kind::tf32: operands rounded to TF32 (round-to-nearest), fp32 accumulate in TMEM */
/* OR-ed into TTAM_PREC_TF32: the operand already holds TF32-representable values (ttam_round_tf32 when it was laid
 * out), so the GEMM skips its in-place rounding pass for it.  X: x of linear_fwd / linear_wgrad; W: w of linear_fwd. */
#define TTAM_PREC_X_ROUNDED 0x100
#define TTAM_PREC_W_ROUNDED 0x200
/* the GEMM writes its result rounded to TF32 (for outputs whose only consumers are tensor-core GEMMs and sign masks:
 * the consumer is then called with TTAM_PREC_X_ROUNDED; the values it multiplies are the same either way) */
#define TTAM_PREC_OUT_ROUNDED 0x400
/* ttam_linear_dgrad only: `w` is the TF32-rounded TRANSPOSED weight [K, N] written by ttam_prepare_weights */
#define TTAM_PREC_WT 0x800
#define TTAM_PREC_BF16 2 /* reserved (the retrieval path, ttam_topk_bf16, is the bf16 tensor-core kernel) */

/* dense optimiser kinds (training.py:1315-1333) */
#define TTAM_OPT_ADAMW 0
#define TTAM_OPT_ADAM 1
#define TTAM_OPT_SGD 2

/* Device-resident per-step state.  Kernels that depend on the step number (bias corrections, dropout
 * counters) read it from here when given a non-null pointer, so that one captured CUDA graph can be
 * replayed for every step; ttam_advance_step bumps it at the head of the step. */
typedef struct {
  int32_t step;        /* 1-based optimiser step of the step in flight */
  int32_t pad_;
  uint64_t rng_offset; /* added to every dropout counter */
} ttam_step_state;
int ttam_advance_step(ttam_step_state* state_dev, uint64_t rng_stride, void* stream);

const char* ttam_last_error(void);
int ttam_version(void);
/* number of kernel launches this library has issued in this process (bench.py's gpu_launches) */
int64_t ttam_launch_count(void);
/* 1 if the library was built with sm_100a SASS and the current device can run it */
int ttam_device_ok(void);

/* ---- row gather ------------------------------------------------------------------------------
 * out[r, 0:ncols] = table[idx[r], 0:ncols].  Replaces nn.Embedding.forward for the ID tables
 * (encoders.py:223), the augmentation tables (adaptive_mimic.py:97-105) and index_select on the
 * feature matrices (training.py:743,747,775). */
int ttam_gather_rows_f32(const float* table, int64_t ld_table, int64_t num_rows, const int64_t* idx,
                         float* out, int64_t ld_out, int64_t R, int64_t ncols, void* stream);
/* in place: x[r, 0:ncols] <- nearest TF32-representable value (10-bit mantissa, ties away from zero: cvt.rna.tf32) */
int ttam_round_tf32(float* x, int64_t ld, int64_t R, int64_t ncols, void* stream);
/* fp32 -> bf16 (round-to-nearest-even) row cast, used to build the retrieval corpus */
int ttam_cast_f32_to_bf16(const float* src, int64_t ld_src, uint16_t* dst, int64_t ld_dst, int64_t R,
                          int64_t ncols, void* stream);

/* ---- linear layers (encoders.py:121-144, 157-162) ------------------------------------------------
 * fwd  : y[M,N] = dropout_p(act(x[gather?][M,K] . w[N,K]^T + bias[N]))      nn.Linear + activation + nn.Dropout
 *        (w rows are ldw floats apart: ldw = K, or a 16-byte-aligned padded copy for the tensor-core path)
 *        if `gather` is non-null the rows of x are x[gather[m]] (fused index_select, training.py:743-775)
 *        dropout: Philox4x32-10 keyed by (seed, offset + state_dev->rng_offset + m*N + n); kept value
 *        scaled by 1/(1-p).  state_dev (nullable) is the device-resident step state below, so that a
 *        captured CUDA graph draws a fresh mask on every replay.
 * dgrad: dx[M,K] (+)= (dy[M,N] . w[N,K]) * act'(...) optional
 *        mask_mode 0: none; 1: multiply by (aux[m,k] > 0)  (ReLU/dropout backward through the saved output)
 *        accumulate != 0 adds into dx
 * wgrad: dw[N,K] = dy[M,N]^T . x[gather?][M,K] ; db[N] = colsum(dy)   (deterministic split-M reduction) */
int ttam_linear_fwd(const float* x, int64_t ldx, const int64_t* gather, const float* w, int64_t ldw, const float* bias,
                    float* y, int64_t ldy, int64_t M, int64_t N, int64_t K, int act, float dropout_p,
                    uint64_t seed, uint64_t offset, const ttam_step_state* state_dev, int precision,
                    void* stream);
int ttam_linear_dgrad(const float* dy, int64_t lddy, const float* w, float* dx, int64_t lddx,
                      const float* aux, int64_t ldaux, int mask_mode, float scale, int accumulate,
                      int64_t M, int64_t N, int64_t K, int precision, void* stream);
/* Once per step, for up to 4 small weight matrices src[i] [rows, cols] (row stride ld): dst[i] = TF32-rounded copy
 * [rows, cols], dst_t[i] = TF32-rounded transposed copy [cols, rows] (either may be null).  With them the forward GEMMs
 * (TTAM_PREC_W_ROUNDED) and the data-gradient GEMMs (TTAM_PREC_WT) of a layer both run on the TMA-fed kernel. */
int ttam_prepare_weights(const float* const* src, const int64_t* ld, const int64_t* rows, const int64_t* cols,
                         float* const* dst, float* const* dst_t, int64_t count, void* stream);
int64_t ttam_linear_wgrad_workspace_bytes(int64_t M, int64_t N, int64_t K);
int ttam_linear_wgrad(const float* dy, int64_t lddy, const float* x, int64_t ldx, const int64_t* gather,
                      float* dw, float* db, int64_t M, int64_t N, int64_t K, int accumulate,
                      void* workspace, int64_t workspace_bytes, int precision, void* stream);

/* elementwise activation + dropout for the non-ReLU feature encoders (encoders.py:68-78,136-137):
 * fwd: y = dropout_p(act(pre));  bwd: dpre = dy * keep/(1-p) * act'(pre)   (mask regenerated from Philox) */
int ttam_act_fwd(const float* pre, float* y, int64_t n, int64_t row_len, int act, float dropout_p, uint64_t seed,
                 uint64_t offset, const ttam_step_state* state_dev, void* stream);
int ttam_act_bwd(const float* dy, const float* pre, float* dpre, int64_t n, int64_t row_len, int act,
                 float dropout_p, uint64_t seed, uint64_t offset, const ttam_step_state* state_dev, void* stream);

/* ---- gated fusion + augmentation (encoders.py:149-168, adaptive_mimic.py:88-105) -----------------
 * z[R,2D] = [e ; f];  pre2[R,D] = G2.relu(G1.z+c1)+c2 (computed by two ttam_linear_fwd calls)
 * fwd: g = sigmoid(pre2); t = g*e + (1-g)*f; o = t + aug[idx]        (aug may be null: o = t)
 * bwd: dpre2 = dt*(e-f)*g*(1-g);  dz = [dt*g ; dt*(1-g)]               */
int ttam_gate_fwd(const float* z, const float* pre2, const float* aug_table, int64_t aug_rows,
                  const int64_t* idx, float* g, float* t, float* o, float* q_out, int64_t R, int64_t D,
                  void* stream);
int ttam_gate_bwd(const float* dt, const float* z, const float* g, float* dpre2, float* dz, int64_t R,
                  int64_t D, void* stream);
/* o[r] = t[r] + aug[idx[r]]   (AdaptiveMimicMechanism._apply_aug, adaptive_mimic.py:88-95) */
int ttam_augment_fwd(const float* t, const float* aug_table, int64_t aug_rows, const int64_t* idx, float* o,
                     float* q_out, int64_t R, int64_t D, void* stream);

/* ---- negative sampling (src/data/samplers.py:11-85) -------------------------------------------------------
 * out[b, n] ~ U[0, num_items) with members of the user's positive set re-drawn, at most 1 + max_rounds draws each
 * (the reference: 10 re-sampling rounds, then RuntimeError).  pos_keys: sorted int64 keys user * num_items + item of
 * every known (user, item) positive (n_keys may be 0).  *fail_flag is set to 1 when a slot is still a positive after
 * the last round (the caller raises).  Draws: Philox4x32-10 keyed by (seed, offset [+ state->rng_offset], element,
 * round): statistical, not bit, parity with torch.randint. */
int ttam_sample_negatives(const int64_t* users, int64_t B, int64_t N, int64_t num_items, const int64_t* pos_keys,
                          int64_t n_keys, int max_rounds, uint64_t seed, uint64_t offset,
                          const ttam_step_state* state_dev, int64_t* out, int32_t* fail_flag, void* stream);

/* ---- composite: one gated tower (1-hidden-layer ReLU MLP) per call -----------------------------------------
 * ttam_tower_fwd / ttam_tower_bwd enqueue the launch sequence of TowerEncoder.forward (encoders.py:221-255) +
 * _apply_aug (adaptive_mimic.py:88-95), resp. their autograd, with ONE C call (an eager step then makes ~10 calls
 * instead of ~85).  Buffers: z [R,2D] = [e ; f], hd [R,H], a [R,Hg], pre2/g/t/o/q [R,D]; gradients: dz [R,2D]
 * (dz[:, :D] = dL/dE rows on return), dpre2 [R,D], dpre1 [R,Hg], dhd [R,H] and the weight gradients. */
typedef struct {
  const float* table;   /* [table_rows, D] ID embedding */
  const float* aug;     /* [table_rows, D] augmentation table, nullable */
  int64_t table_rows, D;
  const float* X;       /* [*, F] feature matrix, rows ldx floats apart */
  int64_t ldx, F;
  const float *W1, *b1; /* [H, F] (rows ldw1 apart), [H] */
  int64_t ldw1, H;
  const float *W2, *b2; /* [D, H], [D] */
  const float *G1, *c1; /* [Hg, 2D], [Hg] */
  int64_t Hg;
  const float *G2, *c2; /* [D, Hg], [D] */
  float dropout_p;
  int32_t precision;
  int32_t x_rounded, w1_rounded; /* X / W1 hold TF32-representable values (see TTAM_PREC_X_ROUNDED) */
  uint64_t seed, rng_base;
  const ttam_step_state* state;
  /* bag form of X (null rowptr: dense layer 1): see ttam_bag_linear_fwd */
  const int64_t* bag_rowptr;
  const void* bag_entries;
  const float* bag_tail;
  int64_t bag_T, bag_tail_start, bag_max_nnz;
  int64_t bag_wgrad;          /* 0: the weight gradient of layer 1 stays the dense GEMM dh^T . X[idx] (rows with dozens of
                                 non-zeros: the scatter costs more than the GEMM), the forward still uses the bag form */
  void* bag_scratch;          /* >= F*H*4 bytes, private to this tower (ttam_tower_fwd transposes W1 into it) */
  int64_t bag_scratch_bytes;
  /* TF32 path: room for the rounded / rounded-transposed copies of W2 [D,H], G1 [Hg,2D], G2 [D,Hg] (all six or none);
   * ttam_tower_fwd fills them (ttam_prepare_weights), forward and backward GEMMs of the chain read them */
  float *W2r, *W2rT, *G1r, *G1rT, *G2r, *G2rT;
} ttam_tower_desc;
typedef struct { float *z, *hd, *a, *pre2, *g, *t, *o, *q; } ttam_tower_bufs;
typedef struct {
  float *dpre2, *dz, *dpre1, *dhd;
  float *dW1, *db1, *dW2, *db2, *dG1, *dc1, *dG2, *dc2;
  int32_t accumulate;
  int32_t phase; /* 0: whole backward; 1: data-gradient chain only (gate_bwd + the three dgrads: what the table updates
                    wait for); 2: the four weight/bias gradients only (may run on another stream once phase 1 is done);
                    10..13: ONE link of the chain (gate_bwd, dgrad through G2, G1, W2); 20..23: ONE weight gradient (G2, G1,
                    W2, W1) - weight gradient 2x needs link 1x only, so the caller can overlap them on two streams */
} ttam_tower_grads;
int ttam_tower_fwd(const ttam_tower_desc* d, const int64_t* idx, int64_t R, const ttam_tower_bufs* bufs, void* stream);
int64_t ttam_tower_bwd_workspace_bytes(const ttam_tower_desc* d, int64_t R);
int ttam_tower_bwd(const ttam_tower_desc* d, const int64_t* idx, int64_t R, const ttam_tower_bufs* bufs, const float* dt,
                   const ttam_tower_grads* grads, void* workspace, int64_t workspace_bytes, void* stream);

/* ---- bag form of the first feature-encoder layer (the "EmbeddingBag" restatement of index_select + nn.Linear(F, H):
 * training.py:743-775 + encoders.py:133; feature layout features.py:242-252) ------------------------------------
 * The feature matrix X [N, F] is held as CSR over its sparse columns [0, tail_start): rowptr int64 [N+1], entries
 * packed 8-byte pairs {int32 column, fp32 value} in column order, at most 64 per row - plus a dense block tail [N, T]
 * (T <= 8, tail_start + T == F) for the trailing columns that are non-zero in most rows (z-scored numerics).
 *   ttam_bag_linear_fwd  : y[r, :] = act(b + sum_j x_j W[:, j]) for x = X[gather[r]] (gather null: row r), fp32 FMA,
 *                          optional Philox dropout (same element numbering as ttam_linear_fwd) and, when
 *                          round_tf32_out != 0, y rounded to TF32 (its consumer is a tensor-core GEMM that skips its
 *                          own rounding pass).  w is the nn.Linear weight [H, F] with row stride ldw.
 *   ttam_bag_linear_wgrad: dw[h, j] (+)= sum_r dh[r, h] x_j(r), db[h] (+)= sum_r dh[r, h].  Deterministic (no atomics).
 * max_nnz = the largest number of CSR entries any row holds (<= 64); it sizes the kernels' shared-memory tile ring.
 * Supported when ttam_bag_supported(H, F, T, max_nnz) != 0 (H % 32 == 0; weight slice + ring fit in shared memory);
 * both take ttam_bag_linear_workspace_bytes(R, H, F) bytes of scratch. */
int ttam_bag_supported(int64_t H, int64_t F, int64_t T, int64_t max_nnz);
int64_t ttam_bag_linear_workspace_bytes(int64_t R, int64_t H, int64_t F);
int ttam_bag_linear_fwd(const int64_t* rowptr, const void* entries, const float* tail, int64_t T, int64_t tail_start,
                        int64_t max_nnz, const int64_t* gather, int64_t R, const float* w, int64_t ldw, const float* bias, float* y,
                        int64_t ldy, int64_t H, int64_t F, int act, float dropout_p, uint64_t seed, uint64_t offset,
                        const void* state_dev, int round_tf32_out, void* workspace, int64_t workspace_bytes, void* stream);
int ttam_bag_linear_wgrad(const int64_t* rowptr, const void* entries, const float* tail, int64_t T, int64_t tail_start,
                          int64_t max_nnz, const int64_t* gather, int64_t R, const float* dh, int64_t lddh, float* dw, int64_t lddw,
                          float* db, int64_t H, int64_t F, int accumulate, void* workspace, int64_t workspace_bytes,
                          void* stream);

/* The same weight gradient on the tensor cores (TF32 products, fp32 accumulation; exact fp32 bias gradient): the CSR rows of a
 * 32-row chunk are expanded into a zeroed shared-memory operand tile and multiplied with tcgen05.mma like a dense x
 * (csrc/gemm_tma.cu bag_wgrad_tc_kernel).  Returns +1 (not an error) when the shape is not covered - H % 32 != 0, F > 640,
 * lddw != F - and the caller should use ttam_bag_linear_wgrad. */
int64_t ttam_bag_linear_wgrad_tc_workspace_bytes(int64_t R, int64_t H, int64_t F);
int ttam_bag_linear_wgrad_tc(const int64_t* rowptr, const void* entries, const float* tail, int64_t T, int64_t tail_start,
                             const int64_t* gather, int64_t R, const float* dh, int64_t lddh, float* dw, int64_t lddw, float* db,
                             int64_t H, int64_t F, int accumulate, void* workspace, int64_t workspace_bytes, void* stream);

/* ---- fused loss forward+backward (training.py:770-803, adaptive_mimic.py:59-68) -------------------
 * o_u[B,D], o_i[(1+N)B,D] (positives first, then negatives row-major [B,N]); t_u,t_p[B,D] base tower
 * outputs, q_u,q_p[B,D] augmentation rows of the positive pairs (all four null when mimic is off).
 * loss_out[4] = {total, bce, mimic_user, mimic_item}.  Gradients (null pointers skip the backward):
 *   do_u[B,D], do_i[(1+N)B,D]  = dL/do = dL/dt ;  dq_u[B,D], dq_p[B,D] = dL/dq of the positive pairs
 *   (dL/dq of a negative row equals its do_i row).
 * batch_fraction (1 on one GPU, 1/world_size when the batch is data-parallel): every mean runs over the GLOBAL
 * batch, so that the per-rank losses and gradients simply add up. */
int64_t ttam_loss_workspace_bytes(int64_t B);
int ttam_loss_fwd_bwd(const float* o_u, const float* o_i, const float* t_u, const float* t_p,
                      const float* q_u, const float* q_p, float lambda_u, float lambda_i, float* loss_out,
                      float* do_u, float* do_i, float* dq_u, float* dq_p, int64_t B, int64_t N, int64_t D,
                      float batch_fraction, void* workspace, int64_t workspace_bytes, void* stream);

/* The same loss with the augmentation add folded in (adaptive_mimic.py:88-95 + training.py:770-803 as ONE pass):
 * t_u[B,D], t_i[(1+N)B,D] base tower outputs (positives first, then negatives [B,N] row-major); aug_u / aug_i the augmentation
 * tables ([rows, D] fp32, rows D floats apart); users[B], items[(1+N)B] the step's row ids (ids outside the table add nothing).
 * o = t + aug[idx] is formed in registers; loss, do_*, dq_* exactly as ttam_augment_fwd followed by ttam_loss_fwd_bwd
 * (mimic != 0: the mimic terms are evaluated, q = the augmentation rows of the positive pairs).  N <= 8, D % 4 == 0, D <= 128
 * (ttam_loss_aug_supported); otherwise use the two calls. */
int ttam_loss_aug_supported(int64_t N, int64_t D);
int ttam_loss_aug_fwd_bwd(const float* t_u, const float* t_i, const float* aug_u, int64_t aug_u_rows, const float* aug_i,
                          int64_t aug_i_rows, const int64_t* users, const int64_t* items, int mimic, float lambda_u,
                          float lambda_i, float* loss_out, float* do_u, float* do_i, float* dq_u, float* dq_p, int64_t B,
                          int64_t N, int64_t D, float batch_fraction, void* workspace, int64_t workspace_bytes, void* stream);

/* The loss of the row-sharded step with the row exchange folded in (SURVEY 8(e) all-to-all #2 / #3 as the loads and stores of ONE
 * kernel; mimic on).  t_u[w] / q_u[w] / t_i[w] / q_i[w] (HOST arrays of `world` device pointers): base of the rows rank w
 * computed for THIS requester's slots (its own buffers through NVLink peer mappings: what ttam_slot_unpack reads);
 * a_*[w] / b_*[w]: base of THIS requester's section of rank w's receive buffers (what ttam_slot_pack writes; rows D floats
 * apart).  slot_of_u[B], slot_of_i[(1+N)B] from ttam_slot_plan (positives first, then negatives [B,N] row-major).
 * Every real slot receives a = do and b = dq (dq of a negative = its do); padding slots are NOT written: the owner zeroes
 * its receive buffers before the step's first barrier.  Same arithmetic and loss_out[4] as ttam_loss_fwd_bwd on
 * o = t + q. */
int ttam_loss_slots_fwd_bwd(const float* const* t_u, const float* const* q_u, const float* const* t_i, const float* const* q_i,
                            float* const* a_u, float* const* b_u, float* const* a_i, float* const* b_i, int64_t world,
                            int64_t cap_u, int64_t cap_i, const int64_t* slot_of_u, const int64_t* slot_of_i, float lambda_u,
                            float lambda_i, float* loss_out, int64_t B, int64_t N, int64_t D, float batch_fraction,
                            void* workspace, int64_t workspace_bytes, void* stream);

/* ---- in-batch softmax loss (EXTENSION: BASELINE.json configs[1] "in-batch negatives"; the reference has no such loss, its
 * training loss is the sampled-negative BCE above - definition and parity: oracle/model.py inbatch_loss_forward_backward) ----
 * S = o_u o_p^T [B,B]; L_ce = mean_b (logsumexp_j S[b,j] - S[b,b]); mimic terms as in ttam_loss_fwd_bwd.
 * loss_out[4] = {total, ce, mimic_user, mimic_item}; do_u, do_p [B,D] = dL/do; dq_u, dq_p [B,D] = dL/dq (mimic only).
 * do_u == NULL: forward only.  precision: TTAM_PREC_FP32 | TTAM_PREC_TF32 (the three B x B x D GEMMs).  The workspace holds S. */
int64_t ttam_inbatch_loss_workspace_bytes(int64_t B, int64_t D);
int ttam_inbatch_loss_fwd_bwd(const float* o_u, const float* o_p, const float* t_u, const float* t_p, const float* q_u,
                              const float* q_p, float lambda_u, float lambda_i, float* loss_out, float* do_u, float* do_p,
                              float* dq_u, float* dq_p, int64_t B, int64_t D, int precision, void* workspace,
                              int64_t workspace_bytes, void* stream);

/* ---- category-alignment loss (training.py:530-579, 805-820) ------------------------------------------------
 * L_cal = mean over the non-major categories c with >= 2 rows of || Cov(emb rows of c) - Cov(emb rows of major) ||_F^2,
 * categories = cat_tensor[item_idx] (primary category id of every item, values in [0, n_categories)).
 * Adds lambda_c * L_cal to loss_out[0] (nullable), writes L_cal to cal_out[0] (nullable) and adds
 * lambda_c * dL_cal/d emb[r] to grad_a[r] (dL/do_i, [R, D]) and, for r < B, to grad_b[r] (dL/dq of the positives,
 * nullable).  Deterministic (sorted segments, fixed-order partial sums), no host synchronisation.  D <= 128. */
int64_t ttam_category_alignment_workspace_bytes(int64_t R, int64_t D, int64_t n_categories);
int ttam_category_alignment(const int64_t* item_idx, int64_t R, const float* emb, int64_t D, const int64_t* cat_tensor,
                            int64_t num_items, int64_t n_categories, int64_t major, float lambda_c, float* loss_out,
                            float* cal_out, float* grad_a, float* grad_b, int64_t B, void* workspace,
                            int64_t workspace_bytes, void* stream);

/* ---- sparse backward + row-wise optimisers ---------------------------------------------------------
 * Step 1  ttam_sort_rows: stable radix sort of the R touched row ids; sorted_idx[R], perm[R] (int32:
 *         original positions).  Replaces coalesce() of the uncoalesced COO gradient
 *         (torch/optim/_functional.py:44, training.py:822).
 * Step 2  one of the row-wise updates; each (unique row) sums its duplicate gradient rows in sorted
 *         (= original) order and applies the optimiser in one pass over p, m, v.  Gradient rows come
 *         from two sources so that no concatenation is materialised: positions < n_a read
 *         grad_a[pos*ld_a], the others grad_b[(pos-n_a)*ld_b].
 *   Hyper-parameters are doubles (Python floats); each is rounded to fp32 where torch rounds it.
 *   ttam_sparse_adam_rows : torch.optim.SparseAdam (torch/optim/_functional.py:24-84); `step` is the
 *                           per-table step count (1-based)
 *   ttam_lazy_rows        : dense AdamW / Adam / SGD(momentum) semantics (torch/optim/adam.py:347-547)
 *                           on a table whose untouched rows are brought up to date lazily:
 *                           last_step[row] records the last step applied; the zero-gradient steps in
 *                           between are replayed in registers before the real one.  scalars[4*t+0] =
 *                           lr/(1-beta1^t), [4*t+1] = sqrt(1-beta2^t), [4*t+2] = lr*sqrt(1-beta2^t)/(1-beta1^t)
 *                           (SparseAdam step size), [4*t+3] = fp32(beta2)^(t/2)/sqrt(1-beta2^t) (zero-gradient replay; see csrc/optim.cu replay()); t = 0..step, computed in double, stored fp32.
 *   Every row-wise update takes `step` by value and, optionally, `state_dev`: when non-null the step is
 *   read from the device (state_dev->step) and the scalars from the table, which makes the launch
 *   replayable inside a CUDA graph.
 *   ttam_lazy_catchup     : replay the touched rows up to step-1 at the HEAD of step `step`, so that the forward
 *                           gather of that step reads what the dense optimiser would have left there.
 *   ttam_lazy_flush       : bring rows [0,num_rows) up to `step` (before any full-table read).
 * ttam_unique_rows: the sorted set of touched rows (what SparseAdam updates), n_unique on the device. */
int64_t ttam_sort_workspace_bytes(int64_t R);
int ttam_sort_rows(const int64_t* idx, int64_t R, int64_t num_rows, int64_t* sorted_idx, int32_t* perm,
                   void* workspace, int64_t workspace_bytes, void* stream);
int ttam_unique_rows(const int64_t* sorted_idx, int64_t R, int64_t* unique_out, int64_t* n_unique_out,
                     void* workspace, int64_t workspace_bytes, void* stream);
/* Long segments: a popular row can own hundreds of the step's gradient rows (Zipf-distributed positives); one warp
 * summing them would serialise the step.  ttam_find_long_segments lists the heads of the segments longer than 16
 * (long_list[0] = count, long_list[1..] = positions in sorted_idx; ttam_long_segments_bytes(R) bytes); a row-wise
 * update that is given the list sums those segments with a whole block each (fixed combination order:
 * deterministic).  long_list may be null: every segment is then summed by one warp. */
int64_t ttam_long_segments_bytes(int64_t R);
int ttam_find_long_segments(const int64_t* sorted_idx, int64_t R, int32_t* long_list, void* stream);
int ttam_sparse_adam_rows(float* p, float* m, float* v, int64_t D, const int64_t* sorted_idx,
                          const int32_t* perm, int64_t R, const float* grad_a, int64_t ld_a, int64_t n_a,
                          const float* grad_b, int64_t ld_b, const float* scalars, double lr, double beta1,
                          double beta2, double eps, int64_t step, const ttam_step_state* state_dev,
                          const int32_t* long_list, void* stream);
int ttam_lazy_rows(int kind, float* p, float* m, float* v, int32_t* last_step, int64_t D,
                   const int64_t* sorted_idx, const int32_t* perm, int64_t R, const float* grad_a,
                   int64_t ld_a, int64_t n_a, const float* grad_b, int64_t ld_b, const float* scalars,
                   double lr, double weight_decay, double beta1, double beta2, double eps, double momentum,
                   int64_t step, const ttam_step_state* state_dev, const int32_t* long_list, void* stream);
/* bring the unique rows of sorted_idx up to step-1 BEFORE the forward pass of step `step` reads them */
int ttam_lazy_catchup(int kind, float* p, float* m, float* v, int32_t* last_step, int64_t D,
                      const int64_t* sorted_idx, int64_t R, const float* scalars, double lr, double weight_decay,
                      double beta1, double beta2, double eps, double momentum, int64_t step,
                      const ttam_step_state* state_dev, void* stream);
int ttam_lazy_flush(int kind, float* p, float* m, float* v, int32_t* last_step, int64_t num_rows, int64_t D,
                    const float* scalars, double lr, double weight_decay, double beta1, double beta2, double eps,
                    double momentum, int64_t step, const ttam_step_state* state_dev, void* stream);
/* dense AdamW / Adam / SGD over up to TTAM_MAX_TENSORS small tensors in ONE launch
 * (the MLP / gate weights and biases; torch/optim/adam.py:347-547) */
#define TTAM_MAX_TENSORS 48
typedef struct {
  int32_t count;
  int32_t pad_;
  float* p[TTAM_MAX_TENSORS];
  const float* g[TTAM_MAX_TENSORS];
  float* m[TTAM_MAX_TENSORS];
  float* v[TTAM_MAX_TENSORS];
  int64_t numel[TTAM_MAX_TENSORS];
} ttam_tensor_list;
int ttam_dense_step(int kind, const ttam_tensor_list* list_host, const float* scalars, double lr,
                    double weight_decay, double beta1, double beta2, double eps, double momentum, int64_t step,
                    const ttam_step_state* state_dev, void* stream);

/* ---- fixed-capacity slot route of the row-sharded step (no reference counterpart: the reference is
 * single-process.  Replaces the per-step split sizes of the id / row / gradient all-to-alls by W x cap static slots,
 * so that the sharded step replays as a CUDA graph.)
 * ttam_slot_plan  : owner(id) = id % world (ids are row ids: >= 0).  Bucket idx[R] by owner, original order kept
 *                   inside a bucket, bucket o
 *                   in slots [o*cap, (o+1)*cap): send_idx[world*cap] (padding slots repeat the bucket's first id: the
 *                   owner's touched-row SET is unchanged), slot_of[R] (slot of every request; world*cap = did not
 *                   fit), req_of[world*cap] (request held by a slot, -1 = padding), *flag = 1 when a bucket holds
 *                   more than cap ids or none (the caller must then route this step dynamically), else 0.
 * ttam_slot_unpack: t_src[w] / q_src[w] (HOST arrays of `world` device pointers, q_src may be NULL) = base of the
 *                   rows owner w produced for THIS requester (row j of bucket w at base + j*ld_src): the local
 *                   receive buffer of an all-to-all, or owner w's own buffer through an NVLink peer mapping.
 *                   Writes, in request order, t_out[R,D], q_out[R,D], o_out[R,D] = t + q (each may be NULL).
 * ttam_slot_pack  : for every slot s (bucket w, position j): a_dst[w][j*ld_dst ..] = a[r], b_dst[w][..] = (r < n0 ?
 *                   b0[r] : b1[r]) with r = req_of[s]; zeros for padding slots.  b_dst may be NULL.
 * ttam_slot_ids   : the id exchange without a collective: src[w] (HOST array of device pointers) = address of the `cap`
 *                   ids requester w bucketed for THIS owner (requester w's send_idx + my_rank*cap, through its peer
 *                   mapping); writes recv_idx[world*cap] and local_rows = recv_idx / world. */
int64_t ttam_slot_plan_workspace_bytes(int64_t R, int64_t world);
int ttam_slot_plan(const int64_t* idx, int64_t R, int64_t world, int64_t cap, int64_t* send_idx, int64_t* slot_of,
                   int32_t* req_of, int32_t* flag, void* workspace, int64_t workspace_bytes, void* stream);
int ttam_slot_unpack(const float* const* t_src, const float* const* q_src, int64_t ld_src, int64_t world, int64_t cap,
                     const int64_t* slot_of, int64_t R, int64_t D, float* t_out, float* q_out, float* o_out,
                     void* stream);
int ttam_slot_ids(const int64_t* const* src, int64_t world, int64_t cap, int64_t* recv_idx, int64_t* local_rows,
                  void* stream);
int ttam_slot_pack(const float* a, const float* b0, int64_t n0, const float* b1, const int32_t* req_of, int64_t world,
                   int64_t cap, int64_t D, float* const* a_dst, float* const* b_dst, int64_t ld_dst, void* stream);

/* ---- retrieval (training.py:330-384, 613-679, 944-972; faiss.IndexFlatIP) --------------------------
 * Exact inner-product top-K of every query against the whole corpus, result in canonical order
 * (descending score, ascending id on ties).  Scores returned are the canonical fp32 scores
 * (sequential accumulation over d).  id_offset is added to every returned id (item-sharded corpora).
 *   ttam_topk_f32 : fp32 operands, SIMT                       (drop-in for the fp32 FAISS path)
 *   ttam_topk_bf16: bf16 operands (D <= 256, K <= 128), TMA-fed tcgen05 GEMM with an in-kernel threshold top-K epilogue,
 *                   followed by an exact re-score + (-score,+id) sort of the survivors
 *   ttam_topk_merge: merge `parts` partial lists per query (item shards) under the same order */
int64_t ttam_topk_f32_workspace_bytes(int64_t Q, int64_t N, int64_t D, int64_t K);
int ttam_topk_f32(const float* q, const float* items, int64_t Q, int64_t N, int64_t D, int64_t K,
                  int64_t id_offset, int64_t* out_ids, float* out_scores, void* workspace,
                  int64_t workspace_bytes, void* stream);
/* Paged form of ttam_topk_f32: query r only admits items that sort strictly AFTER (after_scores[r], after_ids[r]) in the
 * canonical order (after_ids[r] < 0: no restriction), so the caller walks down a ranking K results at a time - what a
 * faiss.IndexFlatIP.search with search_k beyond the kernel's K limit needs (training.py:956-958: search_k grows with the
 * user's number of training positives).  Same workspace as ttam_topk_f32. */
int ttam_topk_f32_after(const float* q, const float* items, int64_t Q, int64_t N, int64_t D, int64_t K,
                        int64_t id_offset, const float* after_scores, const int64_t* after_ids, int64_t* out_ids,
                        float* out_scores, void* workspace, int64_t workspace_bytes, void* stream);
/* out[r, c] = canonical fp32 score of query r against items[cand[r, c]] (cand < 0: -inf): the candidate scoring of
 * `_retrieve_with_sampling` (training.py:986-1005), all users of an evaluation in one launch. */
int ttam_score_pairs(const float* q, const float* items, const int64_t* cand, int64_t R, int64_t C, int64_t D,
                     int64_t N, float* out, void* stream);
int64_t ttam_topk_bf16_workspace_bytes(int64_t Q, int64_t N, int64_t D, int64_t K);
int ttam_topk_bf16(const uint16_t* q, const uint16_t* items, int64_t Q, int64_t N, int64_t D, int64_t K,
                   int64_t id_offset, int64_t* out_ids, float* out_scores, void* workspace,
                   int64_t workspace_bytes, void* stream);
/* fp32 index on the tensor cores (faiss.IndexFlatIP over fp32 embeddings, training.py:646-679, 944-958; K <= 128,
 * D <= 256): every fp32 operand is split as x = hi + lo + r, hi = bf16(x), lo = bf16(x - hi), |r| <= 2^-16 |x|, and laid
 * out as [hi | lo] (each part padded to a multiple of 16 columns: ttam_split_bf16x3_cols(D) = 2 * ceil16(D) columns).  The
 * candidate pass runs the bf16 tcgen05 kernel with THREE products per column from those two copies,
 * qh.xh + ql.xh + qh.xl (hence the name), the survivors within the proven error bound of the K-th score are re-scored from
 * the fp32 rows in the canonical order.  Ids and scores are bit-identical to ttam_topk_f32.  The caller keeps
 * `items_split` next to the corpus (ttam_split_bf16x3 once per index build) and splits each query batch; `item_layout`
 * is ignored (both operands share one layout). */
int64_t ttam_split_bf16x3_cols(int64_t D);
int ttam_split_bf16x3(const float* x, int64_t R, int64_t D, int item_layout, uint16_t* out, void* stream);
int64_t ttam_topk_f32_tc_workspace_bytes(int64_t Q, int64_t N, int64_t D, int64_t K);
int ttam_topk_f32_tc(const float* q, const float* items, const uint16_t* q_split, const uint16_t* items_split,
                     int64_t Q, int64_t N, int64_t D, int64_t K, int64_t id_offset, int64_t* out_ids,
                     float* out_scores, void* workspace, int64_t workspace_bytes, void* stream);
int ttam_topk_merge(const int64_t* ids, const float* scores, int64_t Q, int64_t parts, int64_t K_in,
                    int64_t K_out, int64_t* out_ids, float* out_scores, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* TTAM_H_ */
