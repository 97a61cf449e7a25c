// Full-corpus inner-product top-K on the 5th-gen tensor cores (sm_100a): TMA-fed tcgen05.mma with the score tile
// kept in TMEM and a warp-level running-threshold top-K epilogue; the scores never touch HBM.
// Replaces faiss.IndexFlatIP.search / the chunked torch.topk of the reference evaluation path
// (reference training.py:330-384, 646-679, 944-972) for bf16 corpora (BASELINE.json configs[2]).
//
// Pipeline of ttam_topk_bf16 (all on `stream`, no host synchronisation):
//   1. max_norm_kernel      max ||item||_2 (for the rounding-error bound used in step 3)
//   2. score_topk_kernel    persistent, warp-specialised: 1 TMA warp, 1 MMA warp, 8 epilogue warps per CTA.
//                           Work unit = (block of 256 queries, one of S item ranges).  Per 256-item tile the MMA warp
//                           issues 2 x (D/16) tcgen05.mma (M=128, N=256, K=16, bf16 -> fp32) into two TMEM
//                           accumulators, one per query tile (all 512 TMEM columns; N=256 keeps the operand fetch at
//                           96 B/clk of shared-memory bandwidth - N=128 needs all 128 and ran at half speed).
//                           Epilogue warpgroup a drains accumulator a with tcgen05.ld, releasing it in two 128-column
//                           halves, while the tensor core fills the other one: thread = query row, it owns that
//                           row's running threshold and candidate list.  The hot loop is kept to one copy of the
//                           chunk code (~6 KB) so that it stays in the instruction cache.
//                           A 32-column chunk costs a 3-input-max tree and one compare; the rare survivors are
//                           appended to the row's list in global memory, which a warp-cooperative bitonic sort
//                           compacts to its best 128 whenever it fills up (raising the threshold).
//   3. finalize_kernel      one CTA per query: merge the S lists, take everything within 2*delta of the K-th best
//                           tensor-core score, re-score those candidates in the canonical order (fp32, sequential over
//                           d, the order oracle/retrieval.py defines), sort by (-score, +id), emit K results.  A query
//                           whose candidate set cannot be proven complete is put on the `flagged` list.
//   4. exact_rows_kernel    brute-force canonical scoring for flagged queries only (normally none).
// delta bounds |tensor-core score - canonical score|: both are fp32 accumulations of exact bf16 products.
#include "common.cuh"
#include "sm100.cuh"
#include <cub/cub.cuh>
#include <cuda_bf16.h>
#include <cstdlib>

namespace ttam {
namespace tc {

using namespace ttam::sm100;

constexpr int kBM = 128;       // queries per A tile (UMMA M)
constexpr int kMaxStages = 8;  // ring slots of the item stream (barrier table size)
constexpr uint32_t kMaxSmem = 227 * 1024;
constexpr int kBN = 256;       // items per B tile (one TMA stage)
constexpr int kHN = 128;       // items per MMA (UMMA N): two halves per B tile, one TMEM accumulator each
constexpr int kBoxK = 64;      // bf16 per 128-byte swizzled smem row
constexpr int kCap = 512;      // working list capacity per (query row, item split)
constexpr int kPerLane = kCap / 32;
constexpr int kKeep = 128;     // entries a compaction keeps at least (>= K); exactly this many after the final one
constexpr int kSlack = 32;     // a mid-stream compaction may keep up to kKeep + kSlack (cheaper threshold search)
constexpr int kEpiWarps = 8;
constexpr int kThreads = 32 * (2 + kEpiWarps);
constexpr int kMaxSplits = 16;
constexpr int kMaxCand = 256;  // candidates re-scored per query in the finalize kernel

constexpr uint32_t kABoxBytes = kBM * 128;   // 16 KB: 128 rows x 128 B
constexpr uint32_t kBBoxBytes = kBN * 128;   // 32 KB

// Two shapes of the item ring:
//  * KBOX = 1, 2 (D <= 128; the tuned config-3 path): a stage holds a whole 256-item tile (KBOX boxes of 64 columns), two
//    query tiles (256 queries) per CTA.
//  * KBOX = 0 ("box ring", D up to 768): a stage holds ONE 64-column box of a tile, the MMA thread accumulates over the
//    boxes of a tile as they arrive; p.kbox boxes per tile, p.stages slots - whatever fits beside the resident query tiles.
//    ATILES = 1 (128 queries per CTA, epilogue warpgroup 1 idle) once the query tiles alone would not leave a slot.
//    This is what D = 256 (BASELINE config 4) and the fp32 index (3 x D split-bf16 columns, below) run on.
template <int KBOX>
struct Cfg {
  static constexpr int kStages = KBOX == 1 ? 4 : 2;
  static constexpr uint32_t kABytes = 2 * KBOX * kABoxBytes;
  static constexpr uint32_t kBStage = KBOX * kBBoxBytes;
  static constexpr uint32_t kSmem = kABytes + kStages * kBStage + 1024 /*alignment slack*/ + 256 /*barriers*/;
};

// larger key == better candidate: (score descending, id ascending)
__device__ __forceinline__ uint32_t ordered_bits(float s) {
  if (s == 0.f) s = 0.f;  // -0.0 ties with +0.0
  const uint32_t u = __float_as_uint(s);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float ordered_to_float(uint32_t o) {
  return __uint_as_float((o & 0x80000000u) ? (o & 0x7FFFFFFFu) : ~o);
}
__device__ __forceinline__ uint64_t cand_key(float s, uint32_t id) {
  return ((uint64_t)ordered_bits(s) << 32) | (uint64_t)(0xFFFFFFFFu - id);
}
__device__ __forceinline__ float key_score(uint64_t k) { return ordered_to_float((uint32_t)(k >> 32)); }
__device__ __forceinline__ uint32_t key_id(uint64_t k) { return 0xFFFFFFFFu - (uint32_t)k; }

// ---- list compaction: warp-cooperative threshold selection + in-place filter ------------------------------------------
// List entries are stored raw, {score bits, item id}, so that an append is one 8-byte store.
// compact_row finds t* with  kKeep <= #{score >= t*}  by a bitwise binary search over the order-preserving integer image
// of the scores (one warp reduction per bit, starting below the bits all entries share), keeps the entries >= t*
// (unordered), and raises the row's threshold so that only scores >= t* are appended from now on.  `exact` (the final
// compaction of a work unit) searches down to the last bit and truncates ties so that exactly kKeep entries remain.
// Invariant kept for the finalize kernel: every item of the stream that is NOT in the list scored <= tau.
// Returns {new count, new threshold bits}; meaningful in lane r only.
__device__ __noinline__ uint2 compact_row(int r, int lane, bool exact, uint2* my_buf, int my_cnt, float my_tau) {
  const unsigned long long base = __shfl_sync(0xffffffffu, (unsigned long long)my_buf, r);
  const int n = __shfl_sync(0xffffffffu, my_cnt, r);
  uint2* buf = reinterpret_cast<uint2*>(base);
  __syncwarp();  // lane r's appends are visible to the whole warp
  uint32_t o[kPerLane], id[kPerLane];
  uint32_t lmax = 0u, lmin = 0xFFFFFFFFu;
  // all 16 loads of this lane are issued before the first use (one round trip to L2 instead of sixteen): out-of-range
  // slots read the last valid entry and are masked afterwards
  uint2 raw[kPerLane];
  const int last = n > 0 ? n - 1 : 0;
#pragma unroll
  for (int j = 0; j < kPerLane; ++j) raw[j] = buf[min(j * 32 + lane, last)];
#pragma unroll
  for (int j = 0; j < kPerLane; ++j) {
    const int e = j * 32 + lane;
    const bool ok = e < n;
    o[j] = ok ? ordered_bits(__uint_as_float(raw[j].x)) : 0u;  // 0: below every real entry
    id[j] = ok ? raw[j].y : 0u;
    lmax = max(lmax, o[j]);
    lmin = ok ? min(lmin, o[j]) : lmin;
  }
  const uint32_t omax = __reduce_max_sync(0xffffffffu, lmax);
  const uint32_t omin = __reduce_min_sync(0xffffffffu, lmin);
  uint32_t tstar = omin;  // #{o >= omin} = n >= kKeep
  if (n > kKeep && omax != omin) {
    const int hb = 31 - __clz(omax ^ omin);
    uint32_t prefix = omax & ~((2u << hb) - 1u);  // bits shared by all entries
    for (int bit = hb; bit >= 0; --bit) {
      const uint32_t cand = prefix | (1u << bit);
      int c = 0;
#pragma unroll
      for (int j = 0; j < kPerLane; ++j) c += (o[j] >= cand) ? 1 : 0;
      c = __reduce_add_sync(0xffffffffu, c);
      if (c >= kKeep) {
        prefix = cand;
        if (!exact && c <= kKeep + kSlack) break;
      }
    }
    tstar = prefix;
  }
  // how many are strictly above / tied with t*
  int c_gt = 0, c_eq = 0;
#pragma unroll
  for (int j = 0; j < kPerLane; ++j) {
    c_gt += (o[j] > tstar) ? 1 : 0;
    c_eq += (o[j] == tstar) ? 1 : 0;
  }
  c_gt = __reduce_add_sync(0xffffffffu, c_gt);
  c_eq = __reduce_add_sync(0xffffffffu, c_eq);
  // ties: keep them all when there is room (then scores == t* stay admissible), else only enough to reach kKeep
  const bool keep_all_ties = !exact && (c_gt + c_eq <= kCap - 4 * 64);
  const int tie_budget = keep_all_ties ? c_eq : max(0, min(c_eq, kKeep - c_gt));
  __syncwarp();
  const unsigned lt = (1u << lane) - 1u;
  int total = 0;
#pragma unroll
  for (int j = 0; j < kPerLane; ++j) {
    const bool keep = o[j] > tstar;
    const unsigned bal = __ballot_sync(0xffffffffu, keep);
    if (keep) buf[total + __popc(bal & lt)] = make_uint2(__float_as_uint(ordered_to_float(o[j])), id[j]);
    total += __popc(bal);
  }
  int ties = 0;
#pragma unroll
  for (int j = 0; j < kPerLane; ++j) {
    const bool tie = o[j] == tstar;
    const unsigned bal = __ballot_sync(0xffffffffu, tie);
    const int rank = ties + __popc(bal & lt);
    if (tie && rank < tie_budget) buf[total + rank] = make_uint2(__float_as_uint(ordered_to_float(o[j])), id[j]);
    ties += __popc(bal);
  }
  total += tie_budget;
  __syncwarp();
  if (n > kKeep) {
    my_cnt = total;
    // appended from now on: keep_all_ties -> score >= t*  (tau = the float just below t*), else score > t*
    my_tau = ordered_to_float(keep_all_ties ? tstar - 1u : tstar);
  }
  return make_uint2((uint32_t)my_cnt, __float_as_uint(my_tau));
}

struct MainParams {
  int64_t Q, N;
  int D, S;
  int qblocks, tiles_total, tiles_per_split;
  int sample_tiles;  // T0: tiles of a unit's range visited first in sampling mode (0 = none)
  int kbox, stages;  // box-ring variant only: 64-column boxes per tile, ring slots
  int split_ks;      // box-ring variant, fp32 index: 16-column steps per part of the [hi | lo] operands (0 = plain bf16 operands)
  int trig;          // a row's list is compacted to its best kKeep .. kKeep + kSlack once it holds more than this
  uint2* lists;     // [Q][S][kCap] raw {score bits, id}
  int32_t* cnts;    // [Q][S]
  float* taus;      // [Q][S]
  int debug;        // timing experiments only (TTAM_TOPK_DEBUG): 1 = epilogue drains nothing, 2 = no list appends,
                    // 4 = no TMA item loads after the first stages, 8 = MMA does not wait for the drain
};

// Tile sequence of one work unit (identical in the TMA, MMA and epilogue roles).  The first T0 iterations visit T0 tiles
// spread evenly over the unit's range in SAMPLING mode: the epilogue only tracks, per query row, the 8 largest
// 64-column maxima it sees, and the 8th of them becomes the row's starting threshold.  At least 8 items of the range
// score that high; with a 1/128 sample (61 tiles = 16 k items of 2 M) about a thousand do.  Should fewer than K items of the whole
// corpus beat a row's starting threshold, the finalize kernel sees fewer than K candidates and sends that query to the
// exact fallback, so the result stays exact.  Then every tile of the range is visited in order, normally (the sampled
// tiles a second time: +T0 tiles of tensor work, under 1 %).
// Without this a row starts at -inf, and half of all list appends and most compactions of a stream happen in its
// first 5 %.
struct TileSeq {
  int t0, n_tiles, T0, stride;
  __device__ __forceinline__ TileSeq(const MainParams& p, int s) {
    t0 = s * p.tiles_per_split;
    const int t1 = min(p.tiles_total, t0 + p.tiles_per_split);
    n_tiles = t1 - t0;
    // the 8th largest maximum of a 1/128 sample sits around rank 8 * 128 = 1024 of the range (Gamma(8)-distributed:
    // P(rank < 128) ~ 1e-5), whatever the size of the range.  A larger sample fraction would start tighter but send
    // queries to the exact fallback (measured: 250 k-item shards with 64 sampled tiles -> 780 ms instead of 25 ms).
    T0 = min(p.sample_tiles, n_tiles / 128);
    if (T0 < 4) T0 = 0;
    stride = T0 > 0 ? n_tiles / T0 : 1;
  }
  __device__ __forceinline__ int count() const { return T0 + n_tiles; }
  __device__ __forceinline__ int tile(int i) const { return i < T0 ? t0 + i * stride : t0 + (i - T0); }
};

template <int KBOX, int ATILES>
__global__ void __launch_bounds__(kThreads, 1)
score_topk_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_items, MainParams p) {
  constexpr bool kRing = KBOX == 0;
  constexpr int kQBlock = kBM * ATILES;
  const int kbox = kRing ? p.kbox : KBOX;
  const int n_stages = kRing ? p.stages : (KBOX == 1 ? 4 : 2);
  const uint32_t a_bytes = (uint32_t)(ATILES * kbox) * kABoxBytes;
  const uint32_t stage_bytes = kRing ? kBBoxBytes : (uint32_t)KBOX * kBBoxBytes;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + a_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_b + (uint32_t)n_stages * stage_bytes);
  uint64_t* a_full = bars + 0;
  uint64_t* a_empty = bars + 1;
  uint64_t* acc_full = bars + 2;    // [2] (4 slots reserved): accumulator of query tile a
  uint64_t* acc_empty = bars + 6;   // [4]: half h of accumulator a at index a*2+h
  uint64_t* b_full = bars + 10;     // [kMaxStages]
  uint64_t* b_empty = bars + 10 + kMaxStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 10 + 2 * kMaxStages);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_q);
    tma_prefetch_desc(&tmap_items);
    mbar_init(a_full, 1);
    mbar_init(a_empty, 1);
    for (int i = 0; i < 4; ++i) {
      mbar_init(acc_full + i, 1);
      mbar_init(acc_empty + i, kEpiWarps / 2);  // the 4 warps of the warpgroup that owns query tile a
    }
    for (int i = 0; i < n_stages; ++i) {
      mbar_init(b_full + i, 1);
      mbar_init(b_empty + i, 1);
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int units = p.qblocks * p.S;
  const int ksteps = p.D / 16;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      uint32_t it = 0, un = 0;
      for (int u = blockIdx.x; u < units; u += gridDim.x, ++un) {
        const int qb = u % p.qblocks, s = u / p.qblocks;
        mbar_wait_relaxed(a_empty, (un & 1) ^ 1, 256);
        mbar_expect_tx(a_full, a_bytes);
        for (int a = 0; a < ATILES; ++a)
          for (int kb = 0; kb < kbox; ++kb)
            tma_load_2d(smem_a + (a * kbox + kb) * kABoxBytes, &tmap_q, a_full, kb * kBoxK, qb * kQBlock + a * kBM);
        const TileSeq seq(p, s);
        if constexpr (kRing) {
          // `it` counts boxes here: one ring slot per (tile, 64-column box)
          for (int i = 0; i < seq.count(); ++i) {
            const int t = seq.tile(i);
            for (int kb = 0; kb < kbox; ++kb, ++it) {
              const int stage = it % n_stages;
              mbar_wait_relaxed(b_empty + stage, ((it / n_stages) & 1) ^ 1, 64);
              mbar_expect_tx(b_full + stage, kBBoxBytes);
              tma_load_2d(smem_b + stage * kBBoxBytes, &tmap_items, b_full + stage, kb * kBoxK, t * kBN);
            }
          }
        } else {
          for (int i = 0; i < seq.count(); ++i, ++it) {
            const int t = seq.tile(i);
            const int stage = it % n_stages;
            mbar_wait_relaxed(b_empty + stage, ((it / n_stages) & 1) ^ 1, 256);
            if ((p.debug & 4) && it >= (uint32_t)n_stages) {  // timing experiment: the MMA re-reads stale tiles
              mbar_arrive(b_full + stage);
              continue;
            }
            mbar_expect_tx(b_full + stage, stage_bytes);
            for (int kb = 0; kb < KBOX; ++kb)
              tma_load_2d(smem_b + stage * stage_bytes + kb * kBBoxBytes, &tmap_items, b_full + stage, kb * kBoxK, t * kBN);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (one thread) =====================
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc(1 /*bf16*/, kBM, kBN);
      // descriptors differ only in the 14-bit start-address field: precompute the bases once
      const uint64_t a_desc0 = make_kmajor_desc<128>(smem_u32(smem_a));
      const uint64_t b_desc0 = make_kmajor_desc<128>(smem_u32(smem_b));
      uint32_t it = 0, un = 0, bx = 0;
      for (int u = blockIdx.x; u < units; u += gridDim.x, ++un) {
        const int s = u / p.qblocks;
        const TileSeq seq(p, s);
        mbar_wait(a_full, un & 1);
        if constexpr (kRing) {
          for (int i = 0; i < seq.count(); ++i, ++it) {
            for (int kb = 0; kb < kbox; ++kb, ++bx) {
              const int stage = bx % n_stages;
              mbar_wait(b_full + stage, (bx / n_stages) & 1);
              const uint64_t b_desc = b_desc0 + (uint64_t)((stage * kBBoxBytes) >> 4);
              const int kin = min(4, ksteps - kb * 4);  // 16-column MMA steps inside this box
#pragma unroll 1
              for (int a = 0; a < ATILES; ++a) {
                if (kb == 0) {  // first box of the tile: both halves of accumulator a must have been drained
                  mbar_wait(acc_empty + a * 2, (it & 1) ^ 1);
                  mbar_wait(acc_empty + a * 2 + 1, (it & 1) ^ 1);
                  tc_fence_after();
                }
                const uint32_t d_tmem = tmem_base + (uint32_t)(a * kBN);
                if (p.split_ks == 0) {
                  const uint64_t a_desc = a_desc0 + (uint64_t)(((a * kbox + kb) * kABoxBytes) >> 4);
                  for (int k = 0; k < kin; ++k)
                    umma_f16(d_tmem, a_desc + (uint64_t)(k * 2), b_desc + (uint64_t)(k * 2), idesc, (kb | k) ? 1u : 0u);
                } else {
                  // fp32 index: both operands are [hi | lo], split_ks 16-column steps each.  Item step bk of the hi part meets
                  // the query's hi step bk AND its lo step (qh.xh + ql.xh), a step of the lo part the query's hi step
                  // (qh.xl): three products from two resident copies - the query tiles are addressable at any step.
                  auto a_at = [&](int ak) {
                    return a_desc0 + (uint64_t)((((a * kbox + (ak >> 2)) * kABoxBytes) >> 4) + (uint32_t)((ak & 3) * 2));
                  };
                  for (int k = 0; k < kin; ++k) {
                    const int bk = kb * 4 + k;
                    const uint64_t bd = b_desc + (uint64_t)(k * 2);
                    if (bk < p.split_ks) {
                      umma_f16(d_tmem, a_at(bk), bd, idesc, bk ? 1u : 0u);
                      umma_f16(d_tmem, a_at(p.split_ks + bk), bd, idesc, 1u);
                    } else {
                      umma_f16(d_tmem, a_at(bk - p.split_ks), bd, idesc, 1u);
                    }
                  }
                }
                if (kb == kbox - 1) umma_commit(acc_full + a);
              }
              umma_commit(b_empty + stage);
            }
          }
        } else {
          for (int i = 0; i < seq.count(); ++i, ++it) {
            const int stage = it % n_stages;
            mbar_wait(b_full + stage, (it / n_stages) & 1);
            const uint64_t b_desc = b_desc0 + (uint64_t)((stage * stage_bytes) >> 4);
#pragma unroll 1
            for (int a = 0; a < ATILES; ++a) {
              // one N=256 MMA group fills both 128-column halves of query tile a: both must have been drained
              if (p.debug & 16) {    // (16: timing experiment, polls of the MMA thread back off with nanosleep)
                mbar_wait_relaxed(acc_empty + a * 2, (it & 1) ^ 1, 32);
                mbar_wait_relaxed(acc_empty + a * 2 + 1, (it & 1) ^ 1, 32);
              } else if (!(p.debug & 8)) {  // (8: timing experiment, accumulators overwritten without waiting for the drain)
                mbar_wait(acc_empty + a * 2, (it & 1) ^ 1);
                mbar_wait(acc_empty + a * 2 + 1, (it & 1) ^ 1);
              }
              tc_fence_after();
              const uint32_t d_tmem = tmem_base + (uint32_t)(a * kBN);
              const uint64_t a_desc = a_desc0 + (uint64_t)((a * KBOX * kABoxBytes) >> 4);
              for (int k = 0; k < ksteps; ++k) {
                const uint32_t koff_a = (uint32_t)(((k >> 2) * kABoxBytes + (k & 3) * 32) >> 4);
                const uint32_t koff_b = (uint32_t)(((k >> 2) * kBBoxBytes + (k & 3) * 32) >> 4);
                umma_f16(d_tmem, a_desc + koff_a, b_desc + koff_b, idesc, k > 0 ? 1u : 0u);
              }
              umma_commit(acc_full + a);
            }
            umma_commit(b_empty + stage);
          }
        }
        umma_commit(a_empty);  // all MMAs of this unit have read the query tiles
      }
    }
  } else {
    // ===================== epilogue: warpgroup a (4 warps) owns query tile a; thread = query row ===============
    const int ew = warp - 2;
    const int quad = warp & 3;   // TMEM lanes 32*quad .. +31 are the only ones this warp may read
    const int a = ew >> 2;       // which query tile / pair of accumulators
    const int row_in_tile = quad * 32 + lane;
    uint32_t it = 0;
    for (int u = blockIdx.x; u < (a < ATILES ? units : 0); u += gridDim.x) {
      const int qb = u % p.qblocks, s = u / p.qblocks;
      const TileSeq seq(p, s);
      const int64_t qrow = (int64_t)qb * kQBlock + a * kBM + row_in_tile;
      const bool valid = qrow < p.Q;
      float tau = valid ? -INFINITY : INFINITY;  // rows past the end never collect anything
      int cnt = 0;
      const int trig = p.trig;   // list length that triggers a compaction (<= kCap - 64: room for a chunk pair of appends)
      const int64_t slot = (valid ? qrow : 0) * p.S + s;
      uint2* buf = p.lists + slot * kCap;
      // A pair of 32-column chunks of this thread's row per iteration (two tcgen05.ld in flight, two independent max
      // trees: the loop is latency-bound, not issue-bound): 3-input-max trees, one compare; survivors are rare.
      auto process2 = [&](uint32_t (&v0)[32], uint32_t (&v1)[32], uint32_t id0, bool tail) {
        if (tail) {
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            if ((int64_t)(id0 + i) >= p.N) v0[i] = 0xFF800000u;  // -inf: never a candidate
            if ((int64_t)(id0 + 32 + i) >= p.N) v1[i] = 0xFF800000u;
          }
        }
        float m[16];
#pragma unroll
        for (int g = 0; g < 8; ++g) {
          m[g] = fmaxf(fmaxf(__uint_as_float(v0[4 * g]), __uint_as_float(v0[4 * g + 1])),
                       fmaxf(__uint_as_float(v0[4 * g + 2]), __uint_as_float(v0[4 * g + 3])));
          m[8 + g] = fmaxf(fmaxf(__uint_as_float(v1[4 * g]), __uint_as_float(v1[4 * g + 1])),
                           fmaxf(__uint_as_float(v1[4 * g + 2]), __uint_as_float(v1[4 * g + 3])));
        }
        float m4[4];  // maxima of the four 16-column quarters
#pragma unroll
        for (int q4 = 0; q4 < 4; ++q4) m4[q4] = fmaxf(fmaxf(m[4 * q4], m[4 * q4 + 1]), fmaxf(m[4 * q4 + 2], m[4 * q4 + 3]));
        const float mx = fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3]));
        if (__any_sync(0xffffffffu, mx > tau) && !(p.debug & 2)) {
          // usually one or two of the warp's 2048 scores beat their row's threshold.  Descend into the 16-column quarters
          // and 4-column groups that hold them; the votes of a level are issued back to back BEFORE the first branch on
          // them (a branch between two votes serialises their latencies: the nested any-per-branch form of this descent
          // cost ~40 % of the kernel), the branches are warp-uniform, the appends predicated stores.
          unsigned qv[4];
#pragma unroll
          for (int q4 = 0; q4 < 4; ++q4) qv[q4] = __ballot_sync(0xffffffffu, m4[q4] > tau);
#pragma unroll
          for (int q4 = 0; q4 < 4; ++q4) {
            if (qv[q4]) {
              unsigned gv[4];
#pragma unroll
              for (int gg = 0; gg < 4; ++gg) gv[gg] = __ballot_sync(0xffffffffu, m[4 * q4 + gg] > tau);
#pragma unroll
              for (int gg = 0; gg < 4; ++gg) {
                const int g = 4 * q4 + gg;
                if (gv[gg]) {
#pragma unroll
                  for (int i = 0; i < 4; ++i) {
                    const uint32_t bits = g < 8 ? v0[4 * (g & 7) + i] : v1[4 * (g & 7) + i];
                    if (__uint_as_float(bits) > tau) {
                      buf[cnt] = make_uint2(bits, id0 + 4 * g + i);
                      ++cnt;
                    }
                  }
                }
              }
            }
          }
        }
      };
      auto make_room = [&]() {
        unsigned need = __ballot_sync(0xffffffffu, cnt > trig);
        while (need) {
          const int r = __ffs(need) - 1;
          need &= need - 1;
          const uint2 res = compact_row(r, lane, false, buf, cnt, tau);
          if (lane == r) {
            cnt = (int)res.x;
            tau = __uint_as_float(res.y);
          }
        }
      };
      float top[8];  // sampling mode: the 8 largest 64-column maxima seen so far, descending
#pragma unroll
      for (int j = 0; j < 8; ++j) top[j] = -INFINITY;
      for (int i = 0; i < seq.count(); ++i, ++it) {
        const int t = seq.tile(i);
        const bool sampling = i < seq.T0;
        const bool tail = (t == p.tiles_total - 1) && (p.N % kBN != 0);  // TMA zero-filled rows past the corpus end
        if (p.debug & 8) continue;
        if (i == seq.T0 && seq.T0 > 0 && valid) tau = top[7];  // start of the real pass: adopt the sampled threshold
        mbar_wait(acc_full + a, it & 1);
        tc_fence_after();
#pragma unroll 1
        for (int h = 0; h < 2; ++h) {
          if (!(p.debug & 1)) {
#pragma unroll 1
            for (int cp = 0; cp < 2; ++cp) {
              const int col = h * kHN + cp * 64;
              uint32_t v0[32], v1[32];
              const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(a * kBN + col);
              tmem_ld_32x32_issue(taddr, v0);
              tmem_ld_32x32_issue(taddr + 32u, v1);
              if (!sampling) make_room();
              tmem_ld_wait2(v0, v1);
              if (!sampling) {
                process2(v0, v1, (uint32_t)t * kBN + (uint32_t)col, tail);
              } else {
                const uint32_t id0 = (uint32_t)t * kBN + (uint32_t)col;
                float mx = -INFINITY;
#pragma unroll
                for (int e = 0; e < 32; ++e) {
                  const float x0 = (tail && (int64_t)(id0 + e) >= p.N) ? -INFINITY : __uint_as_float(v0[e]);
                  const float x1 = (tail && (int64_t)(id0 + 32 + e) >= p.N) ? -INFINITY : __uint_as_float(v1[e]);
                  mx = fmaxf(mx, fmaxf(x0, x1));
                }
                if (mx > top[7]) {  // sorted insert: carry the smaller value down
                  float carry = mx;
#pragma unroll
                  for (int j = 0; j < 8; ++j) {
                    const float hi = fmaxf(top[j], carry);
                    carry = fminf(top[j], carry);
                    top[j] = hi;
                  }
                }
              }
            }
          }
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(acc_empty + a * 2 + h);
        }
      }
      // ---- end of unit: leave at most kKeep entries in the list, publish count and threshold
      unsigned need = __ballot_sync(0xffffffffu, cnt > kKeep);
      while (need) {
        const int r = __ffs(need) - 1;
        need &= need - 1;
        const uint2 res = compact_row(r, lane, true, buf, cnt, tau);
        if (lane == r) {
          cnt = (int)res.x;
          tau = __uint_as_float(res.y);
        }
      }
      if (valid) {
        p.cnts[slot] = cnt;
        p.taus[slot] = tau;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}

// ---- max row norm of the corpus -------------------------------------------------------------------------------------
__device__ __forceinline__ float to_f32(float x) { return x; }
__device__ __forceinline__ float to_f32(__nv_bfloat16 x) { return __bfloat162float(x); }

template <typename T>
__global__ void __launch_bounds__(256) max_norm_kernel(const T* __restrict__ items, int64_t N, int D,
                                                       uint32_t* __restrict__ out_bits) {
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  float best = 0.f;
  for (int64_t r = warp0; r < N; r += nwarps) {
    float s = 0.f;
    for (int d = lane; d < D; d += 32) {
      const float x = to_f32(items[r * D + d]);
      s = fmaf(x, x, s);
    }
    s = warp_sum(s);
    best = fmaxf(best, s);
  }
  if (lane == 0) atomicMax(out_bits, __float_as_uint(sqrtf(best)));  // non-negative floats order like their bits
}

// ---- canonical score: fp32, products rounded to fp32 (exact for bf16 inputs), accumulated sequentially over d -------
__device__ __forceinline__ float canonical_dot(const float* __restrict__ qf, const __nv_bfloat16* __restrict__ row, int D) {
  float acc = 0.f;
  for (int d0 = 0; d0 < D; d0 += 8) {
    const uint4 raw = *reinterpret_cast<const uint4*>(row + d0);
    const uint32_t w[4] = {raw.x, raw.y, raw.z, raw.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float lo = __uint_as_float(w[j] << 16), hi = __uint_as_float(w[j] & 0xFFFF0000u);
      acc = __fadd_rn(acc, __fmul_rn(qf[d0 + 2 * j], lo));
      acc = __fadd_rn(acc, __fmul_rn(qf[d0 + 2 * j + 1], hi));
    }
  }
  return acc;
}

__device__ __forceinline__ float canonical_dot(const float* __restrict__ qf, const float* __restrict__ row, int D) {
  float acc = 0.f;
  int d0 = 0;
  if ((reinterpret_cast<uintptr_t>(row) & 15) == 0) {
    for (; d0 + 4 <= D; d0 += 4) {
      const float4 x = *reinterpret_cast<const float4*>(row + d0);
      acc = __fadd_rn(acc, __fmul_rn(qf[d0], x.x));
      acc = __fadd_rn(acc, __fmul_rn(qf[d0 + 1], x.y));
      acc = __fadd_rn(acc, __fmul_rn(qf[d0 + 2], x.z));
      acc = __fadd_rn(acc, __fmul_rn(qf[d0 + 3], x.w));
    }
  }
  for (; d0 < D; ++d0) acc = __fadd_rn(acc, __fmul_rn(qf[d0], row[d0]));
  return acc;
}

// ---- fp32 index on the tensor cores: x = hi + lo + r with hi = bf16(x), lo = bf16(x - hi), |r| <= 2^-16 |x| ------------
// Queries and items are both laid out as [hi | lo] (each part padded to Dp = 16-multiple columns).  The MMA thread forms
// qh.xh + ql.xh + qh.xl = q.x - (ql.xl + residuals) from those two copies (see the split_ks branch of the box-ring issuer):
// |tensor-core score - canonical fp32 score| <= (3.1 * 2^-16 + accumulation terms) * |q| |x|.  The epilogue is the bf16
// kernel's; the finalize kernel re-scores with the fp32 rows.
__global__ void split_bf16_hi_lo_kernel(const float* __restrict__ x, int64_t R, int D, int Dp, __nv_bfloat16* __restrict__ out) {
  const int64_t total = R * Dp;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / Dp;
    const int d = (int)(i - r * Dp);
    const float v = d < D ? x[r * D + d] : 0.f;
    const __nv_bfloat16 hi = __float2bfloat16_rn(v);
    const __nv_bfloat16 lo = __float2bfloat16_rn(v - __bfloat162float(hi));
    __nv_bfloat16* o = out + r * (2 * Dp) + d;
    o[0] = hi;
    o[Dp] = lo;
  }
}

// ascending radix-sort key == (descending score, ascending id)
__device__ __forceinline__ uint64_t final_key(float score, uint32_t id) {
  return ((uint64_t)(~ordered_bits(score)) << 32) | (uint64_t)id;
}
constexpr uint64_t kWorst = ~0ull;

struct FinalParams {
  const __nv_bfloat16* q;      // bf16 index: the operands the tensor cores saw ARE the canonical ones
  const __nv_bfloat16* items;
  const float* qf32;           // fp32 index: canonical operands (the tensor cores saw their [hi | lo] bf16 split)
  const float* itemsf32;
  float delta_rel;             // |tensor-core score - canonical score| <= delta_rel * |q| * max|item|
  int64_t Q, N, id_offset;
  int D, S, K;
  const uint2* lists;
  const int32_t* cnts;
  const float* taus;
  const uint32_t* max_norm_bits;
  int64_t* out_ids;
  float* out_scores;
  int32_t* flagged;      // [Q]
  int32_t* n_flagged;    // [1]
};

template <int ITEMS>
__global__ void __launch_bounds__(256) finalize_kernel(FinalParams p) {
  using SortA = cub::BlockRadixSort<uint64_t, 256, ITEMS>;
  using SortB = cub::BlockRadixSort<uint64_t, 256, 1>;
  __shared__ union {
    typename SortA::TempStorage a;
    typename SortB::TempStorage b;
  } temp;
  __shared__ float qf[256];
  __shared__ uint32_t cand[kMaxCand];
  __shared__ float s_tk, s_qnorm2, s_tq;
  __shared__ int s_total, s_P;
  const int q = blockIdx.x;
  const int tid = threadIdx.x;
  const int L = p.S;
  if (tid < p.D) qf[tid] = p.qf32 ? p.qf32[(int64_t)q * p.D + tid] : __bfloat162float(p.q[(int64_t)q * p.D + tid]);
  if (tid == 0) {
    int tot = 0;
    float tq = -INFINITY;
    for (int l = 0; l < L; ++l) {
      tot += p.cnts[(int64_t)q * L + l];
      tq = fmaxf(tq, p.taus[(int64_t)q * L + l]);
    }
    s_total = tot;
    s_tq = tq;
    s_P = 0;
    s_tk = -INFINITY;
  }
  __syncthreads();
  if (tid == 0) {
    float n2 = 0.f;
    for (int d = 0; d < p.D; ++d) n2 = fmaf(qf[d], qf[d], n2);
    s_qnorm2 = n2;
  }
  uint64_t keys[ITEMS];
#pragma unroll
  for (int i = 0; i < ITEMS; ++i) {
    const int slot = tid * ITEMS + i;
    const int l = slot / kKeep, pos = slot % kKeep;
    uint64_t k = kWorst;
    if (l < L && pos < p.cnts[(int64_t)q * L + l]) {
      const uint2 raw = p.lists[((int64_t)q * L + l) * kCap + pos];
      k = ~cand_key(__uint_as_float(raw.x), raw.y);
    }
    keys[i] = k;
  }
  SortA(temp.a).Sort(keys, 32, 64);  // ascending in ~key over the score bits: best tensor-core score first
  __syncthreads();
  const int total = s_total;
  const int kth = min(p.K, total) - 1;
#pragma unroll
  for (int i = 0; i < ITEMS; ++i)
    if (tid * ITEMS + i == kth) s_tk = key_score(~keys[i]);
  __syncthreads();
  // |tc - canonical| <= delta (bf16 index: both are fp32 accumulations of D exact products bounded by |q||item|)
  const float max_norm = __uint_as_float(*p.max_norm_bits);
  const float delta = p.delta_rel * sqrtf(s_qnorm2) * max_norm + 1e-30f;
  const float thr = s_tk - 2.f * delta;
  int mine = 0;
#pragma unroll
  for (int i = 0; i < ITEMS; ++i)
    if (keys[i] != kWorst && key_score(~keys[i]) >= thr) ++mine;
  if (mine) atomicAdd(&s_P, mine);
  __syncthreads();
  int P = s_P;
  bool flag = false;
  if (P > kMaxCand) {
    P = kMaxCand;
    flag = true;  // too many near-ties to re-score here
  }
  if (s_tq >= thr) flag = true;  // something a list dropped (score <= s_tq) could still belong to the top K
  if (total < p.K) flag = true;  // the sampled starting threshold admitted fewer than K items (never seen in practice)
#pragma unroll
  for (int i = 0; i < ITEMS; ++i) {
    const int rank = tid * ITEMS + i;
    if (rank < P) cand[rank] = key_id(~keys[i]);
  }
  __syncthreads();
  uint64_t fk[1];
  fk[0] = kWorst;
  if (tid < P) {
    const uint32_t id = cand[tid];
    const float sc = p.itemsf32 ? canonical_dot(qf, p.itemsf32 + (int64_t)id * p.D, p.D)
                                : canonical_dot(qf, p.items + (int64_t)id * p.D, p.D);
    fk[0] = final_key(sc, id);
  }
  SortB(temp.b).Sort(fk);
  if (tid < p.K) {
    const bool valid = fk[0] != kWorst;
    p.out_ids[(int64_t)q * p.K + tid] = valid ? (int64_t)(uint32_t)fk[0] + p.id_offset : -1;
    p.out_scores[(int64_t)q * p.K + tid] = valid ? ordered_to_float(~(uint32_t)(fk[0] >> 32)) : -INFINITY;
  }
  if (flag && tid == 0) p.flagged[atomicAdd(p.n_flagged, 1)] = q;
}

// ---- exact brute force for flagged queries (normally zero of them) ---------------------------------------------------
constexpr int kExactItems = 20;
constexpr int kExactChunk = 4096;

__global__ void __launch_bounds__(256) exact_rows_kernel(FinalParams p) {
  using Sort = cub::BlockRadixSort<uint64_t, 256, kExactItems>;
  __shared__ typename Sort::TempStorage temp;
  __shared__ float qf[256];
  const int tid = threadIdx.x;
  const int nf = *p.n_flagged;
  for (int f = blockIdx.x; f < nf; f += gridDim.x) {
    const int q = p.flagged[f];
    __syncthreads();
    if (tid < p.D) qf[tid] = p.qf32 ? p.qf32[(int64_t)q * p.D + tid] : __bfloat162float(p.q[(int64_t)q * p.D + tid]);
    __syncthreads();
    uint64_t keys[kExactItems];
#pragma unroll
    for (int i = 0; i < kExactItems; ++i) keys[i] = kWorst;
    // blocked arrangement after each sort: the running top K lives in slots [0, K) = threads 0 .. ceil(K/20)
    for (int64_t c0 = 0; c0 < p.N; c0 += kExactChunk) {
#pragma unroll
      for (int i = 0; i < kExactItems; ++i) {
        const int slot = tid * kExactItems + i;
        if (slot >= p.K) {
          const int64_t id = c0 + (slot - p.K);
          keys[i] = kWorst;
          if (slot - p.K < kExactChunk && id < p.N)
            keys[i] = final_key(p.itemsf32 ? canonical_dot(qf, p.itemsf32 + id * p.D, p.D) : canonical_dot(qf, p.items + id * p.D, p.D),
                                (uint32_t)id);
        }
      }
      __syncthreads();
      Sort(temp).Sort(keys);
    }
#pragma unroll
    for (int i = 0; i < kExactItems; ++i) {
      const int slot = tid * kExactItems + i;
      if (slot < p.K) {
        const bool valid = keys[i] != kWorst;
        p.out_ids[(int64_t)q * p.K + slot] = valid ? (int64_t)(uint32_t)keys[i] + p.id_offset : -1;
        p.out_scores[(int64_t)q * p.K + slot] = valid ? ordered_to_float(~(uint32_t)(keys[i] >> 32)) : -INFINITY;
      }
    }
  }
}

// ---- host ---------------------------------------------------------------------------------------------------------------
// Item splits per query block.  A split costs more than it saves as soon as the query blocks alone fill the SMs: every
// split runs its own threshold warm-up (the expensive early part of a stream, where most scores still beat the row's
// threshold) and adds a list to merge.  Measured at Q = 100 k x 2 M (391 query blocks, 148 SMs): S = 1: 76.9 ms,
// S = 2: 96.6 ms, S = 3: 106.0 ms, S = 4: 120.0 ms.  Splits therefore only exist to occupy SMs that would otherwise idle.
static int choose_splits(int64_t Q, int64_t N, int qblock) {
  const int64_t qblocks = ceil_div(Q, qblock), tiles = ceil_div(N, kBN);
  const int64_t sms = num_sms();
  int64_t best = 1;
  if (qblocks < sms) {
    best = sms / qblocks;
    if (best > kMaxSplits) best = kMaxSplits;
    const int64_t min_tiles = 64;  // a split shorter than this is all warm-up
    if (best > tiles / min_tiles) best = tiles / min_tiles;
    if (best < 1) best = 1;
  }
  if (const char* e = getenv("TTAM_TOPK_SPLITS")) {
    const int v = atoi(e);
    if (v >= 1 && v <= kMaxSplits && v <= tiles) best = v;
  }
  return (int)best;
}

struct Workspace {
  uint2* lists;
  int32_t* cnts;
  float* taus;
  uint32_t* max_norm_bits;
  int32_t* n_flagged;
  int32_t* flagged;
  int64_t bytes;
};
static Workspace carve(void* base, int64_t Q, int S) {
  Workspace w{};
  int64_t off = 0;
  auto take = [&](int64_t n) {
    void* ptr = base ? (char*)base + off : nullptr;
    off += align_up(n, 256);
    return ptr;
  };
  const int64_t L = Q * S;
  w.lists = (uint2*)take(L * kCap * 8);
  w.cnts = (int32_t*)take(L * 4);
  w.taus = (float*)take(L * 4);
  w.max_norm_bits = (uint32_t*)take(256);
  w.n_flagged = (int32_t*)((char*)w.max_norm_bits + (base ? 128 : 0));
  w.flagged = (int32_t*)take(Q * 4 + 4);
  w.bytes = off + 256;
  return w;
}

// Kernel shape for an MMA depth of Dm bf16 columns (see Cfg): kbox = 0 -> the whole-tile ring (KBOX = 1 or 2).
struct Shape {
  int kbox_t;   // template KBOX (0 = box ring)
  int kbox;     // 64-column boxes per tile
  int atiles;   // query tiles per CTA
  int stages;   // ring slots
  uint32_t smem;
};
static Shape shape_for(int64_t Dm, bool split = false) {
  Shape sh{};
  sh.kbox = (int)ceil_div(Dm, (int64_t)kBoxK);
  const bool force_ring = split || getenv("TTAM_TOPK_RING") != nullptr;   // (env: A/B switch, small D through the box ring)
  if (sh.kbox <= 2 && !force_ring) {
    sh.kbox_t = sh.kbox; sh.atiles = 2;
    sh.stages = sh.kbox == 1 ? 4 : 2;
    sh.smem = sh.kbox == 1 ? Cfg<1>::kSmem : Cfg<2>::kSmem;
    return sh;
  }
  sh.kbox_t = 0;
  sh.atiles = sh.kbox <= 5 ? 2 : 1;
  const int64_t a_bytes = (int64_t)sh.atiles * sh.kbox * kABoxBytes;
  int64_t st = ((int64_t)kMaxSmem - 1024 - 256 - a_bytes) / kBBoxBytes;
  if (st > kMaxStages) st = kMaxStages;
  if (st > 2 * sh.kbox && sh.kbox >= 3) st = 2 * sh.kbox;   // two whole tiles in flight are enough
  sh.stages = (int)st;
  sh.smem = (uint32_t)(a_bytes + st * kBBoxBytes + 1024 + 256);
  return sh;
}

template <int KBOX, int ATILES>
static int launch_main(const CUtensorMap& tq, const CUtensorMap& ti, const MainParams& mp, int grid, uint32_t smem, cudaStream_t st) {
  TTAM_CUDA(cudaFuncSetAttribute(score_topk_kernel<KBOX, ATILES>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  score_topk_kernel<KBOX, ATILES><<<grid, kThreads, smem, st>>>(tq, ti, mp);
  TTAM_LAUNCH_CHECK();
  return TTAM_OK;
}

// The launch sequence shared by the bf16 index (qm / im are the canonical operands) and the fp32 index (qm / im are the
// [hi | lo] bf16 splits of qf / itf, which stay the canonical operands of the re-score).
static int run_topk(const uint16_t* qm, const uint16_t* im, int64_t Dm, const float* qf, const float* itf, int64_t Q,
                    int64_t N, int64_t D, int64_t K, int64_t id_offset, int64_t* out_ids, float* out_scores,
                    void* workspace, int64_t workspace_bytes, cudaStream_t st) {
  const Shape sh = shape_for(Dm, itf != nullptr);
  if (sh.stages < 1) {
    set_error("topk: %lld operand columns do not fit the shared-memory ring", (long long)Dm);
    return TTAM_EUNSUPPORTED;
  }
  const int qblock = kBM * sh.atiles;
  const int S = choose_splits(Q, N, qblock);
  Workspace w = carve(workspace, Q, S);
  if (workspace_bytes < w.bytes) {
    set_error("topk: workspace too small (%lld < %lld)", (long long)workspace_bytes, (long long)w.bytes);
    return TTAM_EWORKSPACE;
  }
  CUtensorMap tq, ti;
  int rc = make_tmap_2d(&tq, qm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, (uint64_t)Q, (uint64_t)Dm, (uint64_t)Dm * 2, kBoxK, kBM,
                        CU_TENSOR_MAP_SWIZZLE_128B);
  if (rc != TTAM_OK) return rc;
  rc = make_tmap_2d(&ti, im, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, (uint64_t)N, (uint64_t)Dm, (uint64_t)Dm * 2, kBoxK, kBN,
                    CU_TENSOR_MAP_SWIZZLE_128B);
  if (rc != TTAM_OK) return rc;

  TTAM_CUDA(cudaMemsetAsync(w.max_norm_bits, 0, 256, st));  // max norm and the flagged counter
  if (itf) max_norm_kernel<float><<<num_sms() * 4, 256, 0, st>>>(itf, N, (int)D, w.max_norm_bits);
  else max_norm_kernel<__nv_bfloat16><<<num_sms() * 4, 256, 0, st>>>((const __nv_bfloat16*)im, N, (int)D, w.max_norm_bits);
  TTAM_LAUNCH_CHECK();

  MainParams mp{};
  mp.Q = Q; mp.N = N; mp.D = (int)Dm; mp.S = S;
  mp.qblocks = (int)ceil_div(Q, qblock);
  mp.tiles_total = (int)ceil_div(N, kBN);
  mp.tiles_per_split = (int)ceil_div(mp.tiles_total, S);
  mp.lists = w.lists; mp.cnts = w.cnts; mp.taus = w.taus;
  mp.kbox = sh.kbox; mp.stages = sh.stages;
  mp.split_ks = itf ? (int)(Dm / 32) : 0;     // Dm = 2 Dp columns = 2 * split_ks steps of 16
  mp.sample_tiles = 64;
  if (const char* e = getenv("TTAM_TOPK_SAMPLE_TILES")) mp.sample_tiles = atoi(e);
  mp.trig = kCap - 64;
  if (const char* e = getenv("TTAM_TOPK_TRIG")) {   // A/B switch
    const int v = atoi(e);
    if (v >= kKeep + kSlack && v <= kCap - 64) mp.trig = v;
  }
  {
    const char* dbg = getenv("TTAM_TOPK_DEBUG");
    mp.debug = dbg ? atoi(dbg) : 0;
  }
  const int units = mp.qblocks * S;
  const int grid = units < num_sms() ? units : num_sms();
  if (sh.kbox_t == 1) rc = launch_main<1, 2>(tq, ti, mp, grid, sh.smem, st);
  else if (sh.kbox_t == 2) rc = launch_main<2, 2>(tq, ti, mp, grid, sh.smem, st);
  else if (sh.atiles == 2) rc = launch_main<0, 2>(tq, ti, mp, grid, sh.smem, st);
  else rc = launch_main<0, 1>(tq, ti, mp, grid, sh.smem, st);
  if (rc != TTAM_OK) return rc;
  if (mp.debug) return TTAM_OK;  // timing experiments: the lists are meaningless, skip the merge

  FinalParams fp{};
  fp.q = (const __nv_bfloat16*)qm; fp.items = (const __nv_bfloat16*)im;
  fp.qf32 = qf; fp.itemsf32 = itf;
  // fp32 accumulation of Dm exact bf16 products on either side (factor 4: alignment truncation inside the tensor core);
  // fp32 index: + the dropped lo.lo products and split residuals (3.1 * 2^-16) + the canonical fp32 sum's own rounding
  fp.delta_rel = 4.f * (float)Dm * 5.9604645e-8f;
  if (itf)   // three products per column: 1.5 * Dm terms in the tensor-core sum
    fp.delta_rel = 1.01f * (1.5f * fp.delta_rel + 3.1f * 1.52587890625e-5f + 2.f * (float)D * 5.9604645e-8f);
  fp.Q = Q; fp.N = N; fp.id_offset = id_offset; fp.D = (int)D; fp.S = S; fp.K = (int)K;
  fp.lists = w.lists; fp.cnts = w.cnts; fp.taus = w.taus; fp.max_norm_bits = w.max_norm_bits;
  fp.out_ids = out_ids; fp.out_scores = out_scores; fp.flagged = w.flagged; fp.n_flagged = w.n_flagged;
  const int per_thread = (int)ceil_div((int64_t)S * kKeep, 256);
  if (per_thread <= 1) finalize_kernel<1><<<(unsigned)Q, 256, 0, st>>>(fp);
  else if (per_thread <= 2) finalize_kernel<2><<<(unsigned)Q, 256, 0, st>>>(fp);
  else if (per_thread <= 4) finalize_kernel<4><<<(unsigned)Q, 256, 0, st>>>(fp);
  else finalize_kernel<8><<<(unsigned)Q, 256, 0, st>>>(fp);
  TTAM_LAUNCH_CHECK();
  exact_rows_kernel<<<num_sms(), 256, 0, st>>>(fp);
  TTAM_LAUNCH_CHECK();
  return TTAM_OK;
}

}  // namespace tc
}  // namespace ttam

using namespace ttam;
using namespace ttam::tc;

constexpr int64_t kMaxD = 256;   // canonical depth (query row held in shared memory by the finalize kernels)

extern "C" int64_t ttam_topk_bf16_workspace_bytes(int64_t Q, int64_t N, int64_t D, int64_t K) {
  (void)K;
  if (Q <= 0 || N <= 0 || D <= 0) return 256;
  return carve(nullptr, Q, choose_splits(Q, N, kBM * shape_for(D).atiles)).bytes;
}

extern "C" int ttam_topk_bf16(const uint16_t* q, const uint16_t* items, int64_t Q, int64_t N, int64_t D, int64_t K,
                              int64_t id_offset, int64_t* out_ids, float* out_scores, void* workspace,
                              int64_t workspace_bytes, void* stream) {
  TTAM_CHECK_ARG(Q >= 0 && N > 0 && D > 0 && K > 0, "topk_bf16: bad shape");
  if (Q == 0) return TTAM_OK;
  TTAM_CHECK_ARG(q && items && out_ids && out_scores && workspace, "topk_bf16: null pointer");
  if (D % 16 != 0 || D > kMaxD || K > kKeep) {
    set_error("topk_bf16: the tcgen05 path needs D %% 16 == 0, D <= %lld and K <= %d (got D=%lld, K=%lld)", (long long)kMaxD,
              kKeep, (long long)D, (long long)K);
    return TTAM_EUNSUPPORTED;
  }
  TTAM_CHECK_ARG(N < (1ll << 32) - 1 && Q < (1ll << 31), "topk_bf16: corpus too large for 32-bit local ids");
  TTAM_CHECK_ARG(((uintptr_t)q & 15) == 0 && ((uintptr_t)items & 15) == 0, "topk_bf16: operands must be 16-byte aligned");
  return run_topk(q, items, D, nullptr, nullptr, Q, N, D, K, id_offset, out_ids, out_scores, workspace, workspace_bytes,
                  (cudaStream_t)stream);
}

// ---- fp32 index on the tensor cores ------------------------------------------------------------------------------------
extern "C" int64_t ttam_split_bf16x3_cols(int64_t D) { return 2 * align_up(D, 16); }

extern "C" int ttam_split_bf16x3(const float* x, int64_t R, int64_t D, int item_layout, uint16_t* out, void* stream) {
  TTAM_CHECK_ARG(R >= 0 && D > 0, "split_bf16x3: bad shape");
  if (R == 0) return TTAM_OK;
  TTAM_CHECK_ARG(x && out, "split_bf16x3: null pointer");
  const int64_t Dp = align_up(D, 16);
  const int64_t blocks = std::min<int64_t>(ceil_div(R * Dp, 256), (int64_t)num_sms() * 16);
  (void)item_layout;   // both operands share the [hi | lo] layout (the argument is kept for ABI stability)
  split_bf16_hi_lo_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(x, R, (int)D, (int)Dp, (__nv_bfloat16*)out);
  TTAM_LAUNCH_CHECK();
  return TTAM_OK;
}

extern "C" int64_t ttam_topk_f32_tc_workspace_bytes(int64_t Q, int64_t N, int64_t D, int64_t K) {
  (void)K;
  if (Q <= 0 || N <= 0 || D <= 0) return 256;
  return carve(nullptr, Q, choose_splits(Q, N, kBM * shape_for(ttam_split_bf16x3_cols(D), true).atiles)).bytes;
}

extern "C" int ttam_topk_f32_tc(const float* q, const float* items, const uint16_t* q_split, const uint16_t* items_split,
                                int64_t Q, int64_t N, int64_t D, int64_t K, int64_t id_offset, int64_t* out_ids,
                                float* out_scores, void* workspace, int64_t workspace_bytes, void* stream) {
  TTAM_CHECK_ARG(Q >= 0 && N > 0 && D > 0 && K > 0, "topk_f32_tc: bad shape");
  if (Q == 0) return TTAM_OK;
  TTAM_CHECK_ARG(q && items && q_split && items_split && out_ids && out_scores && workspace, "topk_f32_tc: null pointer");
  if (D > kMaxD || K > kKeep) {
    set_error("topk_f32_tc: needs D <= %lld and K <= %d (got D=%lld, K=%lld); use ttam_topk_f32", (long long)kMaxD, kKeep,
              (long long)D, (long long)K);
    return TTAM_EUNSUPPORTED;
  }
  TTAM_CHECK_ARG(N < (1ll << 32) - 1 && Q < (1ll << 31), "topk_f32_tc: corpus too large for 32-bit local ids");
  TTAM_CHECK_ARG(((uintptr_t)q_split & 15) == 0 && ((uintptr_t)items_split & 15) == 0,
                 "topk_f32_tc: split operands must be 16-byte aligned");
  return run_topk(q_split, items_split, ttam_split_bf16x3_cols(D), q, items, Q, N, D, K, id_offset, out_ids, out_scores,
                  workspace, workspace_bytes, (cudaStream_t)stream);
}
