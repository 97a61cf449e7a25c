// Bag form of the first feature-encoder layer (the "EmbeddingBag" of BASELINE.json's north star).
//
// The reference gathers dense feature rows x = X[idx] (training.py:743-775; 605 floats, ~4 of them non-zero one-/multi-hot
// columns + a 5-wide numeric tail, features.py:242-252) and multiplies them by a dense weight, nn.Linear(F, H)
// (encoders.py:133).  For such rows  W1 x + b1 = b1 + sum_j x_j W1[:, j]  is a weighted sum over a handful of columns of W1,
// and the weight gradient  dW1 = dh^T X[idx]  is a scatter of  x_j dh_r  rows into the columns j that are non-zero.
// Layout (built once per feature matrix by functional.BagMatrix): CSR over the sparse columns [0, tail_start) - rowptr
// int64 [N+1], entries (int32 column, fp32 value) - plus a dense [N, T] block for the trailing T <= 8 columns that are
// non-zero in most rows (the z-scored numerics).  All arithmetic is fp32 FMA (no TF32 rounding on this layer).
//
//   bag_fwd_kernel   CTA = (slice of SW = 64 | 32 hidden columns, chunk of rows).  The slice of W1^T ([F][SW] fp32, 155 kB at
//                    F = 605) is loaded into shared memory once; then SW/4 lanes per row accumulate b1 + sum_j x_j W1^T[j]
//                    with one 16-byte LDS + 4 FMA per non-zero, apply ReLU / Philox dropout / optional TF32 rounding (when
//                    the consumer is a tensor-core GEMM) and store h with 16-byte stores.  HBM traffic = the output rows.
//   bag_wgrad_kernel same CTA decomposition; a [Fs][SW] fp32 accumulator lives in shared memory.  Rows are staged in tiles
//                    of 32 (dh slice + the rows' entries, cp.async, double-buffered, index chain fetched two tiles ahead).
//                    DETERMINISTIC without atomics: warp w owns the feature columns j with j % 16 == w and applies their
//                    updates in (row, entry) order; the dense tail and the bias gradient accumulate in registers over the
//                    rows a warp owns and are combined in warp order.  Each CTA writes its partial [F+1][SW]; a reduce
//                    kernel sums the chunks in order and stores dW1 in the nn.Linear layout [H][F] (+ db1).
#include "common.cuh"

namespace ttam {
namespace bag {

constexpr int kThreads = 512;
constexpr int kWarps = kThreads / 32;
constexpr int kTileRows = 32;
constexpr int kMaxNnz = 64;   // sparse entries per row (the host builds the bag form only below this)
constexpr int kMaxTail = 8;

struct FwdP {
  const int64_t* rowptr;
  const int2* ent;
  const float* tail;
  int T, tail_start;
  const int64_t* gather;
  int64_t R, rows_per_chunk;
  int F, H;
  const float* WT;  // [F][H]
  const float* bias;
  float* y;
  int64_t ldy;
  int relu;
  float dropout_p;
  uint64_t seed, offset;
  const ttam_step_state* st;
  int round_out;
};

__device__ __forceinline__ float round_tf32(float x) {
  const uint32_t u = __float_as_uint(x);
  return __uint_as_float(((u & 0x7F800000u) != 0x7F800000u) ? ((u + 0x1000u) & 0xFFFFE000u) : u);
}

template <int LPR>
__global__ void __launch_bounds__(kThreads, 1) bag_fwd_kernel(FwdP p) {
  extern __shared__ float4 smem4[];
  constexpr int SW = 4 * LPR;
  constexpr int kRowsPerWarp = 32 / LPR;
  constexpr int kSlots = kThreads / LPR;
  const int h0 = blockIdx.x * SW;
  for (int i = threadIdx.x; i < p.F * LPR; i += kThreads) {
    const int j = i / LPR, c = i - j * LPR;
    smem4[i] = ld_f4(p.WT + (int64_t)j * p.H + h0 + 4 * c);
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int l = threadIdx.x % LPR;
  const int64_t r_begin = (int64_t)blockIdx.y * p.rows_per_chunk;
  const int64_t r_end = min(p.R, r_begin + p.rows_per_chunk);
  const float4 b4 = p.bias ? ld_f4(p.bias + h0 + 4 * l) : make_float4(0.f, 0.f, 0.f, 0.f);
  const bool drop = p.dropout_p > 0.f;
  const float keep_scale = drop ? 1.f / (1.f - p.dropout_p) : 1.f;
  const uint64_t rng_base = p.offset + ((drop && p.st) ? p.st->rng_offset : 0ull);

  // Software pipeline over this row slot's rows (it = 0, 1, ...: row r0 + it * kSlots).  Every load is issued one stage
  // before its consumer so that the dependent chain  gather -> rowptr -> entries  never stalls the in-order issue:
  //   stage 3: g = gather[row(it+3)]      stage 2: (beg, n) = rowptr[g(it+2)]      stage 1: entries / tail of row(it+1)
  constexpr int PASSES = kMaxNnz / LPR;
  const int64_t r0 = r_begin + warp * kRowsPerWarp + lane / LPR;
  auto gidx = [&](int64_t r) -> int64_t { return r < r_end ? (p.gather ? p.gather[r] : r) : -1; };
  auto load_entries = [&](int64_t g, int64_t beg, int n, int2 (&e)[PASSES], float& xt) {
#pragma unroll
    for (int k = 0; k < PASSES; ++k) {
      e[k] = make_int2(0, 0);
      if (k * LPR + l < n) e[k] = p.ent[beg + k * LPR + l];
    }
    xt = (g >= 0 && l < p.T) ? p.tail[g * p.T + l] : 0.f;
  };
  int64_t g_c = gidx(r0), g_1 = gidx(r0 + kSlots), g_2 = gidx(r0 + 2 * (int64_t)kSlots);
  int64_t beg_c = 0, beg_1 = 0;
  int n_c = 0, n_1 = 0;
  if (g_c >= 0) { beg_c = p.rowptr[g_c]; n_c = min((int)(p.rowptr[g_c + 1] - beg_c), kMaxNnz); }
  if (g_1 >= 0) { beg_1 = p.rowptr[g_1]; n_1 = min((int)(p.rowptr[g_1 + 1] - beg_1), kMaxNnz); }
  int2 e_c[PASSES];
  float xt_c;
  load_entries(g_c, beg_c, n_c, e_c, xt_c);
  for (int64_t r = r0; r - lane / LPR < r_end; r += kSlots) {   // warp-uniform condition: the warp's first row
    const bool active = r < r_end;
    // ---- issue the loads of the later stages
    int2 e_n[PASSES];
    float xt_n;
    load_entries(g_1, beg_1, n_1, e_n, xt_n);
    int64_t beg_2 = 0;
    int n_2 = 0;
    if (g_2 >= 0) { beg_2 = p.rowptr[g_2]; n_2 = min((int)(p.rowptr[g_2 + 1] - beg_2), kMaxNnz); }
    const int64_t g_3 = gidx(r + 3 * (int64_t)kSlots);
    // ---- this row
    float4 acc = b4;
    int nmax = active ? n_c : 0;
#pragma unroll
    for (int o = 16; o >= LPR; o >>= 1) nmax = max(nmax, __shfl_xor_sync(0xffffffffu, nmax, o));
#pragma unroll
    for (int k = 0; k < PASSES; ++k) {
      if (k * LPR < nmax) {   // warp-uniform
        const int cnt = min(LPR, nmax - k * LPR);
        for (int i = 0; i < cnt; ++i) {
          const int j = __shfl_sync(0xffffffffu, e_c[k].x, i, LPR);
          const float x = __int_as_float(__shfl_sync(0xffffffffu, e_c[k].y, i, LPR));
          const float4 w = smem4[j * LPR + l];
          acc.x = fmaf(x, w.x, acc.x); acc.y = fmaf(x, w.y, acc.y); acc.z = fmaf(x, w.z, acc.z); acc.w = fmaf(x, w.w, acc.w);
        }
      }
    }
    for (int t = 0; t < p.T; ++t) {
      const float x = __shfl_sync(0xffffffffu, xt_c, t, LPR);
      const float4 w = smem4[(p.tail_start + t) * LPR + l];
      acc.x = fmaf(x, w.x, acc.x); acc.y = fmaf(x, w.y, acc.y); acc.z = fmaf(x, w.z, acc.z); acc.w = fmaf(x, w.w, acc.w);
    }
    if (active) {
      if (p.relu) {
        acc.x = fmaxf(acc.x, 0.f); acc.y = fmaxf(acc.y, 0.f); acc.z = fmaxf(acc.z, 0.f); acc.w = fmaxf(acc.w, 0.f);
      }
      if (drop) {
        // same element numbering as the GEMM epilogue (gemm_tc.cu): element e = r * H + h, keep bits of Philox(seed, e >> 2)
        const uint64_t e0 = rng_base + (uint64_t)r * (uint64_t)p.H + (uint64_t)(h0 + 4 * l);
        if ((e0 & 3) == 0) {   // the usual case (offsets and H are multiples of 4): one Philox block covers the four elements
          const uint4 rb = philox4x32(p.seed, e0 >> 2);
          const float k = 1.0f / 16777216.0f;
          acc.x = (float)(rb.x >> 8) * k >= p.dropout_p ? acc.x * keep_scale : 0.f;
          acc.y = (float)(rb.y >> 8) * k >= p.dropout_p ? acc.y * keep_scale : 0.f;
          acc.z = (float)(rb.z >> 8) * k >= p.dropout_p ? acc.z * keep_scale : 0.f;
          acc.w = (float)(rb.w >> 8) * k >= p.dropout_p ? acc.w * keep_scale : 0.f;
        } else {
          acc.x = dropout_keep(p.seed, e0, p.dropout_p) ? acc.x * keep_scale : 0.f;
          acc.y = dropout_keep(p.seed, e0 + 1, p.dropout_p) ? acc.y * keep_scale : 0.f;
          acc.z = dropout_keep(p.seed, e0 + 2, p.dropout_p) ? acc.z * keep_scale : 0.f;
          acc.w = dropout_keep(p.seed, e0 + 3, p.dropout_p) ? acc.w * keep_scale : 0.f;
        }
      }
      if (p.round_out) {
        acc.x = round_tf32(acc.x); acc.y = round_tf32(acc.y); acc.z = round_tf32(acc.z); acc.w = round_tf32(acc.w);
      }
      st_f4(p.y + r * p.ldy + h0 + 4 * l, acc);
    }
    // ---- rotate the pipeline
#pragma unroll
    for (int k = 0; k < PASSES; ++k) e_c[k] = e_n[k];
    xt_c = xt_n;
    n_c = n_1;
    g_1 = g_2; beg_1 = beg_2; n_1 = n_2;
    g_2 = g_3;
  }
}

// W [H][ldw] -> WT [F][H]
__global__ void transpose_w_kernel(const float* __restrict__ W, int64_t ldw, int H, int F, float* __restrict__ WT) {
  __shared__ float tile[32][33];
  const int j0 = blockIdx.x * 32, h0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int h = h0 + i, j = j0 + threadIdx.x;
    tile[i][threadIdx.x] = (h < H && j < F) ? W[(int64_t)h * ldw + j] : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int j = j0 + i, h = h0 + threadIdx.x;
    if (j < F && h < H) WT[(int64_t)j * H + h] = tile[threadIdx.x][i];
  }
}

// ---- weight gradient -----------------------------------------------------------------------------------------------
struct WgP {
  const int64_t* rowptr;
  const int2* ent;
  const float* tail;
  int T, tail_start;
  const int64_t* gather;
  int64_t R, rows_per_chunk;
  int F, H;
  const float* dh;
  int64_t lddh;
  float* partial;  // [chunks][F + 1][H]
};

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, int valid_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(valid_bytes) : "memory");
}
__device__ __forceinline__ void cp_async8(uint32_t dst, const void* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async4(uint32_t dst, const void* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

struct RowMeta {  // lane i of every warp holds the meta of tile row i
  int64_t g, beg;
  int n, off;
};

template <int SW>
__global__ void __launch_bounds__(kThreads, 1) bag_wgrad_kernel(WgP p) {
  extern __shared__ float4 smem4[];
  constexpr int CPL = SW / 32;  // accumulator columns per lane
  const int Fs = p.tail_start;
  float* acc = reinterpret_cast<float*>(smem4);                         // [Fs][SW]
  float* dh_s = acc + (size_t)Fs * SW;                                  // [2][32][SW]
  int2* ent_s = reinterpret_cast<int2*>(dh_s + 2 * kTileRows * SW);     // [2][32 * kMaxNnz]
  float* tail_s = reinterpret_cast<float*>(ent_s + 2 * kTileRows * kMaxNnz);  // [2][32][8]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int h0 = blockIdx.x * SW;
  const int64_t r_begin = (int64_t)blockIdx.y * p.rows_per_chunk;
  const int64_t r_end = min(p.R, r_begin + p.rows_per_chunk);
  const int ntiles = (int)((r_end - r_begin + kTileRows - 1) / kTileRows);
  for (int i = threadIdx.x; i < Fs * SW / 4; i += kThreads) smem4[i] = make_float4(0.f, 0.f, 0.f, 0.f);

  // the index chain is fetched in two stages (gather three tiles ahead, rowptr two tiles ahead) so that neither load
  // waits for the other inside the loop
  auto load_g = [&](int tile) -> int64_t {
    const int64_t r = r_begin + (int64_t)tile * kTileRows + lane;
    return (tile < ntiles && r < r_end) ? (p.gather ? p.gather[r] : r) : -1;
  };
  auto load_meta = [&](int64_t g) {
    RowMeta m{g, 0, 0, 0};
    if (g >= 0) {
      m.beg = p.rowptr[g];
      m.n = min((int)(p.rowptr[g + 1] - m.beg), kMaxNnz);
    }
    return m;
  };
  auto scan_meta = [&](RowMeta& m) {  // exclusive prefix of n over the tile's rows (every warp computes its own copy)
    int v = m.n;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int u = __shfl_up_sync(0xffffffffu, v, o);
      if (lane >= o) v += u;
    }
    m.off = v - m.n;
  };
  auto stage = [&](int tile, const RowMeta& m) {  // cp.async of tile -> buffer tile & 1
    const int b = tile & 1;
    const int64_t r0 = r_begin + (int64_t)tile * kTileRows;
    // dh slice: 32 rows x SW floats in 16-byte pieces
    for (int i = threadIdx.x; i < kTileRows * (SW / 4); i += kThreads) {
      const int row = i / (SW / 4), c = i - row * (SW / 4);
      const int64_t r = r0 + row;
      const bool ok = r < r_end;
      cp_async16(s32(dh_s + ((size_t)b * kTileRows + row) * SW + 4 * c), ok ? (const void*)(p.dh + r * p.lddh + h0 + 4 * c) : (const void*)p.dh,
                 ok ? 16 : 0);
    }
#pragma unroll
    for (int k = 0; k < 2; ++k) {  // warp w copies the entries / tail of tile rows w and w + 16
      const int row = warp + 16 * k;
      const int64_t beg = __shfl_sync(0xffffffffu, m.beg, row), g = __shfl_sync(0xffffffffu, m.g, row);
      const int n = __shfl_sync(0xffffffffu, m.n, row), off = __shfl_sync(0xffffffffu, m.off, row);
      for (int i = lane; i < n; i += 32) cp_async8(s32(ent_s + (size_t)b * kTileRows * kMaxNnz + off + i), p.ent + beg + i);
      if (lane < p.T && g >= 0) cp_async4(s32(tail_s + ((size_t)b * kTileRows + row) * kMaxTail + lane), p.tail + g * p.T + lane);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };

  float tacc[kMaxTail + 1][CPL];  // dense tail columns + the bias gradient, over the rows this warp owns
#pragma unroll
  for (int t = 0; t <= kMaxTail; ++t)
#pragma unroll
    for (int c = 0; c < CPL; ++c) tacc[t][c] = 0.f;

  RowMeta cur = load_meta(load_g(0));
  scan_meta(cur);
  RowMeta nxt = load_meta(load_g(1));
  int64_t g_ahead = load_g(2);
  __syncthreads();  // accumulator zeroed
  if (ntiles > 0) stage(0, cur);
  for (int tile = 0; tile < ntiles; ++tile) {
    const int b = tile & 1;
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();  // tile `tile` has landed for everyone; everyone is done with tile - 1 (its buffer is free)
    scan_meta(nxt);
    if (tile + 1 < ntiles) stage(tile + 1, nxt);
    RowMeta nn = load_meta(g_ahead);   // rowptr two tiles ahead, gather three tiles ahead: their latency hides behind this tile
    g_ahead = load_g(tile + 3);
    const float* dht = dh_s + (size_t)b * kTileRows * SW;
    const int2* et = ent_s + (size_t)b * kTileRows * kMaxNnz;
    const int rows_here = (int)min((int64_t)kTileRows, r_end - (r_begin + (int64_t)tile * kTileRows));
    for (int i = 0; i < rows_here; ++i) {
      const int n = __shfl_sync(0xffffffffu, cur.n, i);
      if (n == 0) continue;
      const int off = __shfl_sync(0xffffffffu, cur.off, i);
      for (int base = 0; base < n; base += 32) {
        int2 e = make_int2(-1, 0);
        if (base + lane < n) e = et[off + base + lane];
        unsigned mine = __ballot_sync(0xffffffffu, e.x >= 0 && (e.x & (kWarps - 1)) == warp);
        while (mine) {
          const int src = __ffs(mine) - 1;
          mine &= mine - 1;
          const int j = __shfl_sync(0xffffffffu, e.x, src);
          const float x = __int_as_float(__shfl_sync(0xffffffffu, e.y, src));
#pragma unroll
          for (int c = 0; c < CPL; ++c) {
            float* a = acc + (size_t)j * SW + lane + 32 * c;
            *a = fmaf(x, dht[i * SW + lane + 32 * c], *a);
          }
        }
      }
    }
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      const int row = warp + 16 * k;
      if (row < rows_here) {
        float d[CPL];
#pragma unroll
        for (int c = 0; c < CPL; ++c) {
          d[c] = dht[row * SW + lane + 32 * c];
          tacc[kMaxTail][c] += d[c];
        }
#pragma unroll
        for (int t = 0; t < kMaxTail; ++t) {
          if (t < p.T) {
            const float x = tail_s[((size_t)b * kTileRows + row) * kMaxTail + t];
#pragma unroll
            for (int c = 0; c < CPL; ++c) tacc[t][c] = fmaf(x, d[c], tacc[t][c]);
          }
        }
      }
    }
    cur = nxt;
    nxt = nn;
  }
  __syncthreads();
  float* part = p.partial + (size_t)blockIdx.y * (size_t)(p.F + 1) * p.H;
  for (int i = threadIdx.x; i < Fs * (SW / 4); i += kThreads) {
    const int j = i / (SW / 4), c = i - j * (SW / 4);
    st_f4(part + (size_t)j * p.H + h0 + 4 * c, smem4[i]);
  }
  // tail columns and bias: combine the warps' partial sums in warp order
  float* red = dh_s;  // [kWarps][kMaxTail + 1][SW]  (the stage buffers are free now)
#pragma unroll
  for (int t = 0; t <= kMaxTail; ++t)
#pragma unroll
    for (int c = 0; c < CPL; ++c) red[((size_t)warp * (kMaxTail + 1) + t) * SW + lane + 32 * c] = tacc[t][c];
  __syncthreads();
  for (int i = threadIdx.x; i < (kMaxTail + 1) * SW; i += kThreads) {
    const int t = i / SW, col = i - t * SW;
    if (t < p.T || t == kMaxTail) {
      float s = 0.f;
#pragma unroll
      for (int w = 0; w < kWarps; ++w) s += red[((size_t)w * (kMaxTail + 1) + t) * SW + col];
      const int j = t == kMaxTail ? p.F : p.tail_start + t;
      part[(size_t)j * p.H + h0 + col] = s;
    }
  }
}

// dW[h][j] (+)= sum_k partial[k][j][h];  db[h] (+)= sum_k partial[k][F][h]   (chunks summed in order: deterministic)
__global__ void bag_wgrad_reduce_kernel(const float* __restrict__ partial, int chunks, int F, int H, float* __restrict__ dW,
                                        int64_t lddw, float* __restrict__ db, int accumulate) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)(F + 1) * H) return;
  const int j = (int)(i / H), h = (int)(i - (int64_t)j * H);
  float s = 0.f;
  for (int k = 0; k < chunks; ++k) s += partial[((size_t)k * (F + 1) + j) * H + h];
  if (j < F) {
    float* d = dW + (int64_t)h * lddw + j;
    *d = accumulate ? *d + s : s;
  } else if (db) {
    db[h] = accumulate ? db[h] + s : s;
  }
}

static int slice_width(int64_t H, int64_t F) {
  if (H % 64 == 0 && (size_t)F * 64 * 4 <= 172 * 1024) return 64;   // + 50 kB of stage buffers in the wgrad kernel
  if (H % 32 == 0 && (size_t)F * 32 * 4 <= 172 * 1024) return 32;
  return 0;
}
static int64_t chunks_for(int64_t R, int slices) {
  int64_t c = num_sms() / slices;
  if (c < 1) c = 1;
  const int64_t by_rows = ceil_div(R, 64);
  return c < by_rows ? c : (by_rows < 1 ? 1 : by_rows);
}

}  // namespace bag
}  // namespace ttam

using namespace ttam;
using namespace ttam::bag;

extern "C" int ttam_bag_supported(int64_t H, int64_t F, int64_t T) { return slice_width(H, F) != 0 && T >= 0 && T <= kMaxTail && T <= F; }

extern "C" int64_t ttam_bag_linear_workspace_bytes(int64_t R, int64_t H, int64_t F) {
  const int sw = slice_width(H, F);
  if (!sw) return 256;
  const int64_t fwd = align_up(F * H * 4, 256);
  const int64_t wg = chunks_for(R, (int)(H / sw)) * (F + 1) * H * 4;
  return (fwd > wg ? fwd : wg) + 256;
}

extern "C" int ttam_bag_linear_fwd(const int64_t* rowptr, const void* entries, const float* tail, int64_t T, int64_t tail_start,
                                   const int64_t* gather, int64_t R, const float* w, int64_t ldw, const float* bias, float* y,
                                   int64_t ldy, int64_t H, int64_t F, int act, float dropout_p, uint64_t seed, uint64_t offset,
                                   const void* state_dev, int round_tf32_out, void* workspace, int64_t workspace_bytes, void* stream) {
  if (R == 0) return TTAM_OK;
  TTAM_CHECK_ARG(rowptr && entries && w && y && workspace, "bag_linear_fwd: null pointer");
  TTAM_CHECK_ARG(act == TTAM_ACT_NONE || act == TTAM_ACT_RELU, "bag_linear_fwd: fuses ReLU only");
  TTAM_CHECK_ARG(T >= 0 && T <= kMaxTail && tail_start + T == F && (T == 0 || tail), "bag_linear_fwd: bad dense tail");
  const int sw = slice_width(H, F);
  TTAM_CHECK_ARG(sw != 0, "bag_linear_fwd: unsupported shape H=%lld F=%lld (H %% 32 == 0 and F*32*4 <= 176 kB)", (long long)H, (long long)F);
  TTAM_CHECK_ARG(ldy % 4 == 0 && ((uintptr_t)y & 15) == 0 && (!bias || ((uintptr_t)bias & 15) == 0), "bag_linear_fwd: output / bias must be 16-byte aligned");
  if (workspace_bytes < align_up(F * H * 4, 256)) {   // the forward only needs room for W^T
    set_error("bag_linear_fwd: workspace too small");
    return TTAM_EWORKSPACE;
  }
  if (R == 0) return TTAM_OK;
  cudaStream_t s = (cudaStream_t)stream;
  float* WT = (float*)workspace;
  transpose_w_kernel<<<dim3((unsigned)ceil_div(F, 32), (unsigned)ceil_div(H, 32)), dim3(32, 8), 0, s>>>(w, ldw, (int)H, (int)F, WT);
  TTAM_LAUNCH_CHECK();
  FwdP p{};
  p.rowptr = rowptr; p.ent = (const int2*)entries; p.tail = tail; p.T = (int)T; p.tail_start = (int)tail_start; p.gather = gather;
  p.R = R; p.F = (int)F; p.H = (int)H; p.WT = WT; p.bias = bias; p.y = y; p.ldy = ldy; p.relu = act == TTAM_ACT_RELU;
  p.dropout_p = dropout_p; p.seed = seed; p.offset = offset; p.st = (const ttam_step_state*)state_dev; p.round_out = round_tf32_out;
  const int slices = (int)(H / sw);
  const int64_t chunks = chunks_for(R, slices);
  p.rows_per_chunk = ceil_div(R, chunks);
  const size_t smem = (size_t)F * sw * 4;
  dim3 grid((unsigned)slices, (unsigned)chunks);
  if (sw == 64) {
    static bool done = false;
    if (!done) { TTAM_CUDA(cudaFuncSetAttribute(bag_fwd_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024)); done = true; }
    bag_fwd_kernel<16><<<grid, kThreads, smem, s>>>(p);
  } else {
    static bool done = false;
    if (!done) { TTAM_CUDA(cudaFuncSetAttribute(bag_fwd_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024)); done = true; }
    bag_fwd_kernel<8><<<grid, kThreads, smem, s>>>(p);
  }
  TTAM_LAUNCH_CHECK();
  return TTAM_OK;
}

extern "C" int ttam_bag_linear_wgrad(const int64_t* rowptr, const void* entries, const float* tail, int64_t T, int64_t tail_start,
                                     const int64_t* gather, int64_t R, const float* dh, int64_t lddh, float* dw, int64_t lddw,
                                     float* db, int64_t H, int64_t F, int accumulate, void* workspace, int64_t workspace_bytes,
                                     void* stream) {
  TTAM_CHECK_ARG(rowptr && entries && (dh || R == 0) && dw && workspace, "bag_linear_wgrad: null pointer");
  TTAM_CHECK_ARG(T >= 0 && T <= kMaxTail && tail_start + T == F && (T == 0 || tail), "bag_linear_wgrad: bad dense tail");
  const int sw = slice_width(H, F);
  TTAM_CHECK_ARG(sw != 0, "bag_linear_wgrad: unsupported shape H=%lld F=%lld", (long long)H, (long long)F);
  TTAM_CHECK_ARG(lddh % 4 == 0 && ((uintptr_t)dh & 15) == 0, "bag_linear_wgrad: dh must be 16-byte aligned");
  if (workspace_bytes < ttam_bag_linear_workspace_bytes(R, H, F)) {
    set_error("bag_linear_wgrad: workspace too small");
    return TTAM_EWORKSPACE;
  }
  cudaStream_t s = (cudaStream_t)stream;
  const int slices = (int)(H / sw);
  const int64_t chunks = R > 0 ? chunks_for(R, slices) : 0;
  if (R > 0) {
    WgP p{};
    p.rowptr = rowptr; p.ent = (const int2*)entries; p.tail = tail; p.T = (int)T; p.tail_start = (int)tail_start; p.gather = gather;
    p.R = R; p.F = (int)F; p.H = (int)H; p.dh = dh; p.lddh = lddh; p.partial = (float*)workspace;
    p.rows_per_chunk = align_up(ceil_div(R, chunks), kTileRows);
    const size_t smem = ((size_t)tail_start * sw + 2 * kTileRows * sw) * 4 + (size_t)2 * kTileRows * kMaxNnz * 8 +
                        (size_t)2 * kTileRows * kMaxTail * 4;
    const size_t red = (size_t)kWarps * (kMaxTail + 1) * sw * 4;   // aliases the stage buffers: they must be large enough
    const size_t stage_bytes = (size_t)2 * kTileRows * sw * 4 + (size_t)2 * kTileRows * kMaxNnz * 8 + (size_t)2 * kTileRows * kMaxTail * 4;
    TTAM_CHECK_ARG(red <= stage_bytes, "bag_linear_wgrad: internal shared-memory layout");
    TTAM_CHECK_ARG(smem <= 226 * 1024, "bag_linear_wgrad: F too large for the shared-memory accumulator");
    dim3 grid((unsigned)slices, (unsigned)ceil_div(R, p.rows_per_chunk));
    const int64_t real_chunks = grid.y;
    if (sw == 64) {
      static bool done = false;
      if (!done) { TTAM_CUDA(cudaFuncSetAttribute(bag_wgrad_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024)); done = true; }
      bag_wgrad_kernel<64><<<grid, kThreads, smem, s>>>(p);
    } else {
      static bool done = false;
      if (!done) { TTAM_CUDA(cudaFuncSetAttribute(bag_wgrad_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024)); done = true; }
      bag_wgrad_kernel<32><<<grid, kThreads, smem, s>>>(p);
    }
    TTAM_LAUNCH_CHECK();
    bag_wgrad_reduce_kernel<<<(unsigned)ceil_div((F + 1) * H, 256), 256, 0, s>>>((const float*)workspace, (int)real_chunks, (int)F, (int)H, dw, lddw,
                                                                                  db, accumulate);
    TTAM_LAUNCH_CHECK();
  } else if (!accumulate) {
    bag_wgrad_reduce_kernel<<<(unsigned)ceil_div((F + 1) * H, 256), 256, 0, s>>>((const float*)workspace, 0, (int)F, (int)H, dw, lddw, db, 0);
    TTAM_LAUNCH_CHECK();
  }
  return TTAM_OK;
}
