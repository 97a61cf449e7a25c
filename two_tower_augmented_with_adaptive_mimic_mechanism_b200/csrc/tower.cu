// Composite entry points: the whole forward / backward of one gated tower (1-hidden-layer ReLU MLP) as ONE C call that
// enqueues the same launch sequence tower_ops.py issues op by op (reference encoders.py:221-255 + adaptive_mimic.py:88-95
// and their autograd).  Nothing new is computed here; the point is host-side: an eager step (the row-sharded N-GPU
// step, the training hooks) makes ~10 C calls instead of ~85, so the CPU stays ahead of the GPU.
#include "common.cuh"
#include <algorithm>

using namespace ttam;

#define TTAM_TRY(expr)            \
  do {                            \
    const int rc__ = (expr);      \
    if (rc__ != TTAM_OK) return rc__; \
  } while (0)

extern "C" int ttam_tower_fwd(const ttam_tower_desc* d, const int64_t* idx, int64_t R, const ttam_tower_bufs* b, void* stream) {
  TTAM_CHECK_ARG(d && b && (R == 0 || idx), "tower_fwd: null pointer");
  TTAM_CHECK_ARG(d->table && (d->X || d->bag_rowptr) && d->W1 && d->W2 && d->G1 && d->G2, "tower_fwd: incomplete tower description");
  TTAM_CHECK_ARG(b->z && b->hd && b->a && b->pre2 && b->g && b->t, "tower_fwd: missing activation buffer");
  if (R == 0) return TTAM_OK;
  const int64_t D = d->D, H = d->H, Hg = d->Hg, F = d->F;
  // e = E[idx] -> z[:, :D]
  TTAM_TRY(ttam_gather_rows_f32(d->table, D, d->table_rows, idx, b->z, 2 * D, R, D, stream));
  // hd = dropout(relu(X[idx] W1^T + b1))
  const bool bag = d->bag_rowptr != nullptr;
  const bool tc = (d->precision & 0xFF) != TTAM_PREC_FP32;
  const bool prep = tc && d->W2r && d->W2rT && d->G1r && d->G1rT && d->G2r && d->G2rT;
  if (prep) {   // rounded + rounded-transposed copies of the three small weights, one launch; the backward reuses them
    const float* src[3] = {d->W2, d->G1, d->G2};
    const int64_t ld[3] = {H, 2 * D, Hg}, rows[3] = {D, Hg, D}, cols[3] = {H, 2 * D, Hg};
    float* dst[3] = {d->W2r, d->G1r, d->G2r};
    float* dst_t[3] = {d->W2rT, d->G1rT, d->G2rT};
    TTAM_TRY(ttam_prepare_weights(src, ld, rows, cols, dst, dst_t, 3, stream));
  }
  const int wr = prep ? TTAM_PREC_W_ROUNDED : 0;
  if (bag) {
    // bag form: b1 + sum_j x_j W1[:, j] in fp32, written TF32-rounded when the next GEMM runs on the tensor cores
    TTAM_CHECK_ARG(d->bag_scratch, "tower_fwd: the bag form needs bag_scratch (room for W1^T)");
    TTAM_TRY(ttam_bag_linear_fwd(d->bag_rowptr, d->bag_entries, d->bag_tail, d->bag_T, d->bag_tail_start, d->bag_max_nnz, idx, R, d->W1, d->ldw1,
                                 d->b1, b->hd, H, H, F, TTAM_ACT_RELU, d->dropout_p, d->seed, d->rng_base, d->state, tc ? 1 : 0,
                                 d->bag_scratch, d->bag_scratch_bytes, stream));
  } else {
    const int prec1 = d->precision | (d->x_rounded ? TTAM_PREC_X_ROUNDED : 0) | (d->w1_rounded ? TTAM_PREC_W_ROUNDED : 0);
    TTAM_TRY(ttam_linear_fwd(d->X, d->ldx, idx, d->W1, d->ldw1, d->b1, b->hd, H, R, H, F, TTAM_ACT_RELU, d->dropout_p, d->seed,
                             d->rng_base, d->state, prec1, stream));
  }
  // f = hd W2^T + b2 -> z[:, D:]
  TTAM_TRY(ttam_linear_fwd(b->hd, H, nullptr, prep ? d->W2r : d->W2, H, d->b2, b->z + D, 2 * D, R, D, H, TTAM_ACT_NONE, 0.f, 0, 0, nullptr,
                           d->precision | wr | ((bag && tc) ? TTAM_PREC_X_ROUNDED : 0), stream));
  // a = relu(z G1^T + c1);  pre2 = a G2^T + c2.  With prepared weights `a` is written TF32-rounded: its consumers are the
  // next GEMM, the ReLU mask of the backward and the weight-gradient GEMM - the values they use are the same either way
  TTAM_TRY(ttam_linear_fwd(b->z, 2 * D, nullptr, prep ? d->G1r : d->G1, 2 * D, d->c1, b->a, Hg, R, Hg, 2 * D, TTAM_ACT_RELU, 0.f, 0, 0,
                           nullptr, d->precision | wr | (prep ? TTAM_PREC_OUT_ROUNDED : 0), stream));
  TTAM_TRY(ttam_linear_fwd(b->a, Hg, nullptr, prep ? d->G2r : d->G2, Hg, d->c2, b->pre2, D, R, D, Hg, TTAM_ACT_NONE, 0.f, 0, 0, nullptr,
                           d->precision | wr | (prep ? TTAM_PREC_X_ROUNDED : 0), stream));
  // g = sigmoid(pre2); t = g e + (1-g) f; o = t + A[idx]
  return ttam_gate_fwd(b->z, b->pre2, d->aug, d->aug ? d->table_rows : 0, idx, b->g, b->t, d->aug ? b->o : nullptr,
                       d->aug ? b->q : nullptr, R, D, stream);
}

extern "C" int64_t ttam_tower_bwd_workspace_bytes(const ttam_tower_desc* d, int64_t R) {
  if (!d || R <= 0) return 256;
  int64_t m = ttam_linear_wgrad_workspace_bytes(R, d->D, d->Hg);
  const int64_t c[3] = {ttam_linear_wgrad_workspace_bytes(R, d->Hg, 2 * d->D), ttam_linear_wgrad_workspace_bytes(R, d->D, d->H),
                        (d->bag_rowptr && d->bag_wgrad) ? std::max(ttam_bag_linear_workspace_bytes(R, d->H, d->F), ttam_bag_linear_wgrad_tc_workspace_bytes(R, d->H, d->F))
                                                        : ttam_linear_wgrad_workspace_bytes(R, d->H, d->F)};
  for (int i = 0; i < 3; ++i) m = c[i] > m ? c[i] : m;
  return m;
}

extern "C" int ttam_tower_bwd(const ttam_tower_desc* d, const int64_t* idx, int64_t R, const ttam_tower_bufs* b, const float* dt,
                              const ttam_tower_grads* g, void* workspace, int64_t workspace_bytes, void* stream) {
  TTAM_CHECK_ARG(d && b && g && dt && workspace && (R == 0 || idx), "tower_bwd: null pointer");
  TTAM_CHECK_ARG(g->dpre2 && g->dz && g->dpre1 && g->dhd && g->dW1 && g->db1 && g->dW2 && g->db2 && g->dG1 && g->dc1 && g->dG2 && g->dc2,
                 "tower_bwd: missing gradient buffer");
  if (R == 0) return TTAM_OK;
  const int64_t D = d->D, H = d->H, Hg = d->Hg, F = d->F;
  const int acc = g->accumulate;
  const int prec = d->precision;
  TTAM_CHECK_ARG((g->phase >= 0 && g->phase <= 2) || (g->phase >= 10 && g->phase <= 13) || (g->phase >= 20 && g->phase <= 23),
                 "tower_bwd: phase must be 0, 1, 2, 10..13 or 20..23");
  // 0: everything; 1: the data-gradient chain; 2: the four weight gradients; 10..13 / 20..23: ONE link of the chain / ONE
  // weight gradient (gate, gate layer 2, gate layer 1, MLP layer 2 | G2, G1, W2, W1), so that the caller can start each
  // weight gradient on a second stream as soon as the chain link that produces its input has run
  const int ph = g->phase;
  const bool chain = ph != 2 && ph < 20, weights = ph != 1 && (ph < 10 || ph >= 20);
  const auto link = [&](int i) { return ph < 10 || ph == 10 + i; };
  const auto wg = [&](int i) { return ph < 10 || ph == 20 + i; };
  const float* df = g->dz + D;   // feature MLP: df = dz[:, D:]
  const float scale = d->dropout_p > 0.f ? 1.f / (1.f - d->dropout_p) : 1.f;
  const bool prep = (prec & 0xFF) != TTAM_PREC_FP32 && d->W2r && d->W2rT && d->G1r && d->G1rT && d->G2r && d->G2rT;
  if (chain) {
    // dpre2 = dt (e-f) g (1-g);  dz = [dt g ; dt (1-g)]
    if (link(0)) TTAM_TRY(ttam_gate_bwd(dt, b->z, b->g, g->dpre2, g->dz, R, D, stream));
    if (prep) {   // data gradients against the transposed weight copies the forward prepared: TMA-fed kernel
      if (link(1)) TTAM_TRY(ttam_linear_dgrad(g->dpre2, D, d->G2rT, g->dpre1, Hg, b->a, Hg, 1, 1.f, 0, R, D, Hg, prec | TTAM_PREC_WT | TTAM_PREC_OUT_ROUNDED, stream));
      if (link(2)) TTAM_TRY(ttam_linear_dgrad(g->dpre1, Hg, d->G1rT, g->dz, 2 * D, nullptr, 0, 0, 1.f, 1, R, Hg, 2 * D, prec | TTAM_PREC_WT | TTAM_PREC_X_ROUNDED, stream));
      if (link(3)) TTAM_TRY(ttam_linear_dgrad(df, 2 * D, d->W2rT, g->dhd, H, b->hd, H, 1, scale, 0, R, D, H, prec | TTAM_PREC_WT, stream));
    } else {
      if (link(1)) TTAM_TRY(ttam_linear_dgrad(g->dpre2, D, d->G2, g->dpre1, Hg, b->a, Hg, 1, 1.f, 0, R, D, Hg, prec, stream));
      if (link(2)) TTAM_TRY(ttam_linear_dgrad(g->dpre1, Hg, d->G1, g->dz, 2 * D, nullptr, 0, 0, 1.f, 1, R, Hg, 2 * D, prec, stream));
      if (link(3)) TTAM_TRY(ttam_linear_dgrad(df, 2 * D, d->W2, g->dhd, H, b->hd, H, 1, scale, 0, R, D, H, prec, stream));
    }
  }
  if (weights) {
    if (wg(0)) TTAM_TRY(ttam_linear_wgrad(g->dpre2, D, b->a, Hg, nullptr, g->dG2, g->dc2, R, D, Hg, acc, workspace, workspace_bytes,
                                          prec | (prep ? TTAM_PREC_X_ROUNDED : 0), stream));
    if (wg(1)) TTAM_TRY(ttam_linear_wgrad(g->dpre1, Hg, b->z, 2 * D, nullptr, g->dG1, g->dc1, R, Hg, 2 * D, acc, workspace, workspace_bytes, prec, stream));
    const bool bag = d->bag_rowptr != nullptr;
    const bool tc = (prec & 0xFF) != TTAM_PREC_FP32;
    if (wg(2)) TTAM_TRY(ttam_linear_wgrad(df, 2 * D, b->hd, H, nullptr, g->dW2, g->db2, R, D, H, acc, workspace, workspace_bytes,
                                          prec | ((bag && tc) ? TTAM_PREC_X_ROUNDED : 0), stream));
    if (!wg(3)) {
    } else if (bag && d->bag_wgrad) {
      // tensor-core path: the CSR rows expanded into the MMA operand tile (TF32, like the other weight gradients of the step);
      // +1 = shape not covered -> the fp32 SIMT scatter kernel
      int rc = 1;
      if (tc && workspace_bytes >= ttam_bag_linear_wgrad_tc_workspace_bytes(R, H, F))
        rc = ttam_bag_linear_wgrad_tc(d->bag_rowptr, d->bag_entries, d->bag_tail, d->bag_T, d->bag_tail_start, idx, R, g->dhd, H, g->dW1, F,
                                      g->db1, H, F, acc, workspace, workspace_bytes, stream);
      if (rc < 0) return rc;
      if (rc == 1)
      TTAM_TRY(ttam_bag_linear_wgrad(d->bag_rowptr, d->bag_entries, d->bag_tail, d->bag_T, d->bag_tail_start, d->bag_max_nnz, idx, R, g->dhd, H,
                                     g->dW1, F, g->db1, H, F, acc, workspace, workspace_bytes, stream));
    } else {
      TTAM_TRY(ttam_linear_wgrad(g->dhd, H, d->X, d->ldx, idx, g->dW1, g->db1, R, H, F, acc, workspace, workspace_bytes,
                                 prec | (d->x_rounded ? TTAM_PREC_X_ROUNDED : 0), stream));
    }
  }
  return TTAM_OK;
}
