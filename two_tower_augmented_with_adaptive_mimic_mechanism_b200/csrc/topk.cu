// Exact inner-product top-K (fp32 SIMT path), candidate re-scoring and shard merge.
// Replaces faiss.IndexFlatIP.search + host filtering input (reference training.py:672-675,955-958) and the
// chunked torch.topk of _score_all_items_for_user (training.py:354-382).
// Result order is canonical: descending score, ascending id on ties.
#include "gemm_simt.cuh"
#include <cub/cub.cuh>

namespace ttam {

// ---- order-preserving key: ascending uint64 == (descending score, ascending id) -----------------------------
__host__ __device__ __forceinline__ uint64_t make_key(float score, uint32_t id) {
  if (score == 0.f) score = 0.f;  // -0.0 -> +0.0 so that it ties with +0.0
  uint32_t u;
#ifdef __CUDA_ARCH__
  u = __float_as_uint(score);
#else
  memcpy(&u, &score, 4);
#endif
  uint32_t ordered = (u & 0x80000000u) ? ~u : (u | 0x80000000u);  // ascending with the float order
  return ((uint64_t)(~ordered) << 32) | (uint64_t)id;
}
__device__ __forceinline__ float key_score(uint64_t key) {
  uint32_t ordered = ~(uint32_t)(key >> 32);
  uint32_t u = (ordered & 0x80000000u) ? (ordered & 0x7FFFFFFFu) : ~ordered;
  return __uint_as_float(u);
}
constexpr uint64_t kWorstKey = ~0ull;  // sorts last; decodes to id 0xFFFFFFFF

constexpr int kSelThreads = 256;
constexpr int kSelItems = 20;                       // 5120 keys per block
constexpr int kSelCap = kSelThreads * kSelItems;

constexpr int kChunk = 4096;                        // items per scoring pass
constexpr int kPend = kSelCap - 1024;               // pending candidates per query (>= kChunk; K + kPend <= kSelCap)
static_assert(kPend >= kChunk && kChunk % kSelThreads == 0, "a chunk of survivors must fit the pending list");

// running <- best K of (running  U  pending[0, cnt)); blocked arrangement: thread t owns slots t*kSelItems .. +kSelItems-1
__device__ __forceinline__ void merge_pending(uint64_t* __restrict__ running, const uint64_t* __restrict__ pend, int cnt,
                                              int K, void* temp_storage) {
  using Sort = cub::BlockRadixSort<uint64_t, kSelThreads, kSelItems>;
  uint64_t keys[kSelItems];
#pragma unroll
  for (int i = 0; i < kSelItems; ++i) {
    const int slot = threadIdx.x * kSelItems + i;
    uint64_t k = kWorstKey;
    if (slot < K) k = running[slot];
    else if (slot - K < cnt) k = pend[slot - K];
    keys[i] = k;
  }
  Sort(*reinterpret_cast<typename Sort::TempStorage*>(temp_storage)).Sort(keys);
#pragma unroll
  for (int i = 0; i < kSelItems; ++i) {
    const int slot = threadIdx.x * kSelItems + i;
    if (slot < K) running[slot] = keys[i];
  }
}

// One block per query.  A chunk score becomes a candidate only if it beats the query's current K-th best (the last key
// of `running`) - after the first chunks that is a handful of the 4096 - and candidates wait in `pending` until that
// fills up; only then is the block-wide sort paid (K ln(N / K) candidates in all: two or three sorts per query over a
// whole corpus instead of one per chunk).
// scores: [Q, ld] fp32 for ids id0 .. id0+n-1.  running: [Q, K] keys (sorted ascending).  pending: [Q, kPend], cnt [Q].
// after: optional [Q] keys; a candidate is admitted only if it sorts strictly AFTER its query's key (paged search).
__global__ void __launch_bounds__(kSelThreads) select_topk_kernel(const float* __restrict__ scores, int64_t ld, int n,
                                                                  uint32_t id0, uint64_t* __restrict__ running, int K,
                                                                  uint64_t* __restrict__ pending, int32_t* __restrict__ pcnt,
                                                                  const uint64_t* __restrict__ after) {
  using Sort = cub::BlockRadixSort<uint64_t, kSelThreads, kSelItems>;
  __shared__ typename Sort::TempStorage temp;
  __shared__ int s_new, s_base;
  const int q = blockIdx.x;
  uint64_t* run = running + (int64_t)q * K;
  uint64_t* pend = pending + (int64_t)q * kPend;
  const uint64_t floor_key = after ? after[q] : 0ull;
  constexpr int kPer = kChunk / kSelThreads;   // scores per thread (n <= kChunk), strided: coalesced loads
  uint64_t thr = run[K - 1];
  if (threadIdx.x == 0) s_new = 0;
  __syncthreads();
  uint64_t keys[kPer];
  int mine = 0;
#pragma unroll
  for (int i = 0; i < kPer; ++i) {
    const int j = i * kSelThreads + threadIdx.x;
    uint64_t k = kWorstKey;
    if (j < n) {
      k = make_key(scores[(int64_t)q * ld + j], id0 + (uint32_t)j);
      if (k >= thr || (after && k <= floor_key)) k = kWorstKey;
    }
    keys[i] = k;
    mine += k != kWorstKey;
  }
  int at = mine ? atomicAdd(&s_new, mine) : 0;
  __syncthreads();
  const int total = s_new;
  if (total == 0) return;
  int cnt = pcnt[q];
  if (cnt + total > kPend) {   // block-uniform: fold the pending candidates into `running`, tighten the threshold, re-filter
    __syncthreads();
    merge_pending(run, pend, cnt, K, &temp);
    if (threadIdx.x == 0) s_new = 0;
    __syncthreads();
    thr = run[K - 1];
    mine = 0;
#pragma unroll
    for (int i = 0; i < kPer; ++i) {
      if (keys[i] >= thr) keys[i] = kWorstKey;
      mine += keys[i] != kWorstKey;
    }
    at = mine ? atomicAdd(&s_new, mine) : 0;
    __syncthreads();
    cnt = 0;
  }
#pragma unroll
  for (int i = 0; i < kPer; ++i)
    if (keys[i] != kWorstKey) pend[cnt + at++] = keys[i];
  __syncthreads();
  if (threadIdx.x == 0) pcnt[q] = cnt + s_new;
}

// end of the corpus: fold what is still pending
__global__ void __launch_bounds__(kSelThreads) merge_pending_kernel(uint64_t* __restrict__ running, int K,
                                                                    const uint64_t* __restrict__ pending,
                                                                    const int32_t* __restrict__ pcnt) {
  using Sort = cub::BlockRadixSort<uint64_t, kSelThreads, kSelItems>;
  __shared__ typename Sort::TempStorage temp;
  const int q = blockIdx.x;
  const int cnt = pcnt[q];
  if (cnt == 0) return;
  merge_pending(running + (int64_t)q * K, pending + (int64_t)q * kPend, cnt, K, &temp);
}

// (score, returned id) of the last result of the previous page -> key; id < 0: no previous page (admit everything)
__global__ void after_keys_kernel(const float* __restrict__ scores, const int64_t* __restrict__ ids, int64_t id_offset,
                                  int64_t Q, uint64_t* __restrict__ keys) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < Q) keys[i] = ids[i] < 0 ? 0ull : make_key(scores[i], (uint32_t)(ids[i] - id_offset));
}

// out[r, c] = <q[r, :], items[cand[r, c], :]> with the canonical arithmetic (products rounded to fp32, accumulated in
// the order d = 0..D-1); cand < 0 -> -inf.  One thread per pair: candidate lists are short (reference training.py:974-1009:
// ground truth + 50 sampled items per user).
__global__ void score_pairs_kernel(const float* __restrict__ q, const float* __restrict__ items, const int64_t* __restrict__ cand,
                                   int64_t R, int64_t C, int D, int64_t N, float* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= R * C) return;
  const int64_t r = i / C, id = cand[i];
  if (id < 0 || id >= N) {
    out[i] = -INFINITY;
    return;
  }
  const float* a = q + r * D;
  const float* b = items + id * D;
  float acc = 0.f;
  for (int d = 0; d < D; ++d) acc = __fadd_rn(acc, __fmul_rn(a[d], b[d]));
  out[i] = acc;
}

__global__ void fill_keys_kernel(uint64_t* keys, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    keys[i] = kWorstKey;
}

__global__ void decode_keys_kernel(const uint64_t* __restrict__ keys, int64_t n, int64_t id_offset,
                                   int64_t* __restrict__ ids, float* __restrict__ scores) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const uint64_t k = keys[i];
    if (k == kWorstKey) {
      ids[i] = -1;
      if (scores) scores[i] = -INFINITY;
    } else {
      ids[i] = (int64_t)(uint32_t)k + id_offset;
      if (scores) scores[i] = key_score(k);
    }
  }
}

// merge `parts` sorted/unsorted lists of K_in (score,id) pairs per query into the best K_out (ids are global int64,
// so the key carries the position and the id is looked up afterwards; ties broken by the int64 id).
constexpr int kMergeThreads = 128;
constexpr int kMergeItems = 8;  // 1024 candidates per query
struct MergeKey {
  float score;
  int64_t id;
};
__global__ void __launch_bounds__(kMergeThreads) merge_kernel(const int64_t* __restrict__ ids,
                                                              const float* __restrict__ scores, int total, int K_out,
                                                              int64_t* __restrict__ out_ids,
                                                              float* __restrict__ out_scores) {
  // two-pass radix sort: stable sort by id ascending, then stable sort by score descending
  using SortId = cub::BlockRadixSort<int64_t, kMergeThreads, kMergeItems, float>;
  using SortSc = cub::BlockRadixSort<uint32_t, kMergeThreads, kMergeItems, int64_t>;
  __shared__ union {
    typename SortId::TempStorage a;
    typename SortSc::TempStorage b;
  } temp;
  const int q = blockIdx.x;
  int64_t kid[kMergeItems];
  float ksc[kMergeItems];
#pragma unroll
  for (int i = 0; i < kMergeItems; ++i) {
    const int slot = threadIdx.x * kMergeItems + i;
    if (slot < total && ids[(int64_t)q * total + slot] >= 0) {
      kid[i] = ids[(int64_t)q * total + slot];
      ksc[i] = scores[(int64_t)q * total + slot];
    } else {
      kid[i] = INT64_MAX;
      ksc[i] = -INFINITY;
    }
  }
  SortId(temp.a).Sort(kid, ksc);
  __syncthreads();
  uint32_t skey[kMergeItems];
#pragma unroll
  for (int i = 0; i < kMergeItems; ++i) skey[i] = (uint32_t)(make_key(ksc[i], 0) >> 32);
  SortSc(temp.b).Sort(skey, kid);  // LSD radix sort is stable: equal scores keep ascending id order
#pragma unroll
  for (int i = 0; i < kMergeItems; ++i) {
    const int slot = threadIdx.x * kMergeItems + i;
    if (slot < K_out) {
      const bool valid = kid[i] != INT64_MAX;
      out_ids[(int64_t)q * K_out + slot] = valid ? kid[i] : -1;
      out_scores[(int64_t)q * K_out + slot] = valid ? key_score((uint64_t)skey[i] << 32) : -INFINITY;
    }
  }
}

constexpr int kQBlock = 2048;   // queries per scoring pass

}  // namespace ttam

using namespace ttam;

extern "C" int64_t ttam_topk_f32_workspace_bytes(int64_t Q, int64_t N, int64_t D, int64_t K) {
  (void)N; (void)D;
  int64_t qb = Q < kQBlock ? Q : kQBlock;
  return align_up(qb * kChunk * 4, 256) + align_up(Q * K * 8, 256) + align_up(Q * 8, 256) +
         align_up(qb * (int64_t)kPend * 8, 256) + align_up(qb * 4, 256) + 256;
}

extern "C" int ttam_topk_f32(const float* q, const float* items, int64_t Q, int64_t N, int64_t D, int64_t K,
                             int64_t id_offset, int64_t* out_ids, float* out_scores, void* workspace,
                             int64_t workspace_bytes, void* stream) {
  return ttam_topk_f32_after(q, items, Q, N, D, K, id_offset, nullptr, nullptr, out_ids, out_scores, workspace,
                             workspace_bytes, stream);
}

extern "C" int ttam_score_pairs(const float* q, const float* items, const int64_t* cand, int64_t R, int64_t C, int64_t D,
                                int64_t N, float* out, void* stream) {
  TTAM_CHECK_ARG(q && items && cand && out, "score_pairs: null pointer");
  TTAM_CHECK_ARG(R >= 0 && C >= 0 && D > 0 && N > 0, "score_pairs: bad shape");
  if (R * C == 0) return TTAM_OK;
  score_pairs_kernel<<<(unsigned)ceil_div(R * C, 256), 256, 0, (cudaStream_t)stream>>>(q, items, cand, R, C, (int)D, N, out);
  TTAM_LAUNCH_CHECK();
  return TTAM_OK;
}

extern "C" int ttam_topk_f32_after(const float* q, const float* items, int64_t Q, int64_t N, int64_t D, int64_t K,
                                   int64_t id_offset, const float* after_scores, const int64_t* after_ids,
                                   int64_t* out_ids, float* out_scores, void* workspace, int64_t workspace_bytes,
                                   void* stream) {
  TTAM_CHECK_ARG((after_scores == nullptr) == (after_ids == nullptr), "topk_f32_after: pass both after arrays or neither");
  TTAM_CHECK_ARG(q && items && out_ids && workspace, "topk_f32: null pointer");
  TTAM_CHECK_ARG(Q >= 0 && N > 0 && D > 0 && K > 0, "topk_f32: bad shape");
  TTAM_CHECK_ARG(K + kChunk <= kSelCap, "topk_f32: K must be <= %d", kSelCap - kChunk);
  TTAM_CHECK_ARG(N < (1ll << 32) - 1, "topk_f32: corpus too large for 32-bit local ids");
  if (workspace_bytes < ttam_topk_f32_workspace_bytes(Q, N, D, K)) {
    set_error("topk_f32: workspace too small");
    return TTAM_EWORKSPACE;
  }
  if (Q == 0) return TTAM_OK;
  cudaStream_t s = (cudaStream_t)stream;
  const int64_t qb = Q < kQBlock ? Q : kQBlock;
  float* scores = (float*)workspace;
  uint64_t* running = (uint64_t*)((char*)workspace + align_up(qb * kChunk * 4, 256));
  uint64_t* after_buf = running + align_up(Q * K * 8, 256) / 8;
  uint64_t* pending = after_buf + align_up(Q * 8, 256) / 8;
  int32_t* pcnt = (int32_t*)(pending + align_up(qb * (int64_t)kPend * 8, 256) / 8);
  uint64_t* after = nullptr;
  if (after_ids) {
    after = after_buf;
    after_keys_kernel<<<(unsigned)ceil_div(Q, 256), 256, 0, s>>>(after_scores, after_ids, id_offset, Q, after);
    TTAM_LAUNCH_CHECK();
  }
  fill_keys_kernel<<<(int)std::min<int64_t>(ceil_div(Q * K, 256), 4096), 256, 0, s>>>(running, Q * K);
  TTAM_LAUNCH_CHECK();
  for (int64_t q0 = 0; q0 < Q; q0 += qb) {
    const int64_t nq = std::min<int64_t>(qb, Q - q0);
    TTAM_CUDA(cudaMemsetAsync(pcnt, 0, (size_t)nq * 4, s));
    for (int64_t c0 = 0; c0 < N; c0 += kChunk) {
      const int64_t nc = std::min<int64_t>(kChunk, N - c0);
      GemmP p{};
      p.A = q + q0 * D; p.B = items + c0 * D; p.C = scores; p.lda = D; p.ldb = D; p.ldc = kChunk;
      p.M = (int)nq; p.N = (int)nc; p.K = (int)D; p.scale = 1.f;
      dim3 grid((unsigned)ceil_div(nc, BN), (unsigned)ceil_div(nq, BM), 1);
      gemm_f32_kernel<true, true, true><<<grid, 256, 0, s>>>(p);
      TTAM_LAUNCH_CHECK();
      select_topk_kernel<<<(unsigned)nq, kSelThreads, 0, s>>>(scores, kChunk, (int)nc, (uint32_t)c0, running + q0 * K,
                                                              (int)K, pending, pcnt, after ? after + q0 : nullptr);
      TTAM_LAUNCH_CHECK();
    }
    merge_pending_kernel<<<(unsigned)nq, kSelThreads, 0, s>>>(running + q0 * K, (int)K, pending, pcnt);
    TTAM_LAUNCH_CHECK();
  }
  decode_keys_kernel<<<(int)std::min<int64_t>(ceil_div(Q * K, 256), 4096), 256, 0, s>>>(running, Q * K, id_offset,
                                                                                        out_ids, out_scores);
  TTAM_LAUNCH_CHECK();
  return TTAM_OK;
}

extern "C" int ttam_topk_merge(const int64_t* ids, const float* scores, int64_t Q, int64_t parts, int64_t K_in,
                               int64_t K_out, int64_t* out_ids, float* out_scores, void* stream) {
  TTAM_CHECK_ARG(ids && scores && out_ids && out_scores, "topk_merge: null pointer");
  TTAM_CHECK_ARG(parts > 0 && K_in > 0 && K_out > 0 && parts * K_in <= kMergeThreads * kMergeItems &&
                     K_out <= parts * K_in,
                 "topk_merge: parts*K_in must be <= %d and K_out <= parts*K_in", kMergeThreads * kMergeItems);
  if (Q == 0) return TTAM_OK;
  merge_kernel<<<(unsigned)Q, kMergeThreads, 0, (cudaStream_t)stream>>>(ids, scores, (int)(parts * K_in), (int)K_out,
                                                                        out_ids, out_scores);
  TTAM_LAUNCH_CHECK();
  return TTAM_OK;
}
