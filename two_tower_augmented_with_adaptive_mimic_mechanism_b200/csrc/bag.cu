// Bag form of the first feature-encoder layer (the "EmbeddingBag" of BASELINE.json's north star).
//
// The reference gathers dense feature rows x = X[idx] (training.py:743-775; 605 floats, ~4 of them non-zero one-/multi-hot
// columns + a 5-wide numeric tail, features.py:242-252) and multiplies them by a dense weight, nn.Linear(F, H)
// (encoders.py:133).  For such rows  W1 x + b1 = b1 + sum_j x_j W1[:, j]  is a weighted sum over a handful of columns of W1,
// and the weight gradient  dW1 = dh^T X[idx]  is a scatter of  x_j dh_r  rows into the columns j that are non-zero.
// Layout (built once per feature matrix by functional.BagMatrix): CSR over the sparse columns [0, tail_start) - rowptr
// int64 [N+1], entries (int32 column, fp32 value), at most max_nnz <= 64 per row - plus a dense [N, T] block for the
// trailing T <= 8 columns that are non-zero in most rows (the z-scored numerics).  All arithmetic is fp32 FMA.
//
// Both kernels: CTA = (slice of SW = 64 | 32 hidden columns, chunk of <= 1024 rows), 16 warps, one CTA per SM.
//   1. the dependent index chain  gather[r] -> rowptr[g]  of the WHOLE chunk is resolved up front by all 512 threads at once
//      (two global-memory latencies per CTA instead of two per tile) into a 12-byte-per-row table in shared memory
//      (row id, first entry, entry count + offset inside the row's tile);
//   2. the rows are then staged through shared memory in tiles of TR = 64 | 32 rows by a 3-stage cp.async ring (two tiles in
//      flight while one is processed): entries, dense tail (and, for the weight gradient, the dh slice).
//   bag_fwd_kernel   the slice of W1^T ([F][SW] fp32, 155 kB at F = 605) is loaded into shared memory once; SW/4 lanes per
//                    row accumulate b1 + sum_j x_j W1^T[j] (entry = one broadcast 8-byte LDS, weight = one 16-byte LDS,
//                    4 FMA), apply ReLU / Philox dropout / optional TF32 rounding (when the consumer is a tensor-core GEMM)
//                    and store h with 16-byte stores.  HBM traffic = the CSR rows + the output rows.
//   bag_wgrad_kernel a [Fs][SW] fp32 accumulator lives in shared memory; the ring also carries the dh slice of the tile.
//                    DETERMINISTIC without atomics: warp w owns the feature columns j with j % 16 == w; every warp scans
//                    the tile's flat entry list 32 entries at a time and applies the updates of its columns in (row,
//                    entry) order.  The dense tail and the bias gradient accumulate in registers over the rows a warp owns
//                    and are combined in warp order.  Each CTA writes its partial [F+1][SW]; a reduce kernel sums the
//                    chunks in order and stores dW1 in the nn.Linear layout [H][F] (+ db1).
#include "common.cuh"

namespace ttam {
namespace bag {

constexpr int kThreads = 512;
constexpr int kWarps = kThreads / 32;
constexpr int kStages = 3;
constexpr int kMaxNnz = 64;   // sparse entries per row (the host builds the bag form only below this)
constexpr int kMaxTail = 8;
constexpr size_t kSmemLimit = 226 * 1024;

struct BagP {
  const int64_t* rowptr;
  const int2* ent;
  const float* tail;
  int T, tail_start;
  const int64_t* gather;
  int64_t R, rows_per_chunk;
  int F, H;
  int TR, ecap;  // tile rows (32 | 64), entry capacity of a stage (>= TR * max_nnz)
  // forward
  const float* WT;  // [F][H]
  const float* bias;
  float* y;
  int64_t ldy;
  int relu;
  float dropout_p;
  uint64_t seed, offset;
  const ttam_step_state* st;
  int round_out;
  // weight gradient
  const float* dh;
  int64_t lddh;
  float* partial;  // [chunks][F + 1][H]
};

__device__ __forceinline__ float round_tf32(float x) {
  const uint32_t u = __float_as_uint(x);
  return __uint_as_float(((u & 0x7F800000u) != 0x7F800000u) ? ((u + 0x1000u) & 0xFFFFE000u) : u);
}
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

// ---- shared-memory layout of the ring (identical on host and device) ---------------------------------------------
struct Layout {
  size_t main_bytes;   // W slice (fwd) / accumulator (wgrad)
  size_t ent_off, tail_off, dh_off, rowid_off, stage_bytes;
  size_t meta_off, meta_bytes, total;
};
constexpr int kChunkRows = 1024;  // rows per CTA (a multiple of every TR): bounds the meta table
__host__ __device__ inline Layout make_layout(size_t main_floats, int TR, int ecap, int SW, bool wgrad) {
  Layout L;
  L.main_bytes = (main_floats * 4 + 127) / 128 * 128;
  size_t o = 0;
  L.ent_off = o; o += (size_t)ecap * 8;
  L.tail_off = o; o += (size_t)TR * kMaxTail * 4;
  L.dh_off = o; o += wgrad ? (size_t)TR * SW * 4 : 0;
  L.rowid_off = o; o += wgrad ? ((size_t)ecap + 15) / 16 * 16 : 0;
  L.stage_bytes = (o + 127) / 128 * 128;
  L.meta_off = L.main_bytes + kStages * L.stage_bytes;
  // meta table of the chunk: g[rows] (int32), beg[rows] (uint32), offn[rows] (uint32: offset in the tile << 8 | count)
  L.meta_bytes = (size_t)kChunkRows * 12;
  L.total = L.meta_off + L.meta_bytes;
  return L;
}

struct MetaView {   // the chunk's row table; row = index inside the chunk
  int* g;
  uint32_t* beg;
  uint32_t* offn;
  __device__ __forceinline__ int n(int row) const { return (int)(offn[row] & 0xFFu); }
  __device__ __forceinline__ int off(int row) const { return (int)(offn[row] >> 8); }
};

// The tile pipeline shared by both kernels: run<ROWID>(extra, process) with extra(tile, stage) issuing additional
// cp.async copies of a tile (wgrad: its dh slice) and process(tile, stage, meta, first row of the tile in the chunk).
struct Ring {
  const BagP& p;
  uint8_t* smem;
  Layout L;
  int64_t chunk_begin, chunk_end;  // this CTA's rows
  int64_t r_begin, r_end;          // the block of <= kChunkRows rows of the chunk that is in the table / ring right now
  int ntiles, lane, warp, rpl;
  int pad_mask;  // 3: every row's entry list is padded to a multiple of 4 with zero entries (forward), 0: compact (wgrad)
  MetaView m;

  __device__ __forceinline__ void set_block(int64_t first) {
    r_begin = chunk_begin + first;
    r_end = min(chunk_end, r_begin + kChunkRows);
    ntiles = (int)((r_end - r_begin + p.TR - 1) / p.TR);
  }

  __device__ Ring(const BagP& p_, uint8_t* smem_, const Layout& L_, int pad_mask_) : p(p_), smem(smem_), L(L_), pad_mask(pad_mask_) {
    lane = threadIdx.x & 31;
    warp = threadIdx.x >> 5;
    rpl = p.TR / 32;
    chunk_begin = (int64_t)blockIdx.y * p.rows_per_chunk;
    chunk_end = min(p.R, chunk_begin + p.rows_per_chunk);
    set_block(0);
    m.g = reinterpret_cast<int*>(smem + L.meta_off);
    m.beg = reinterpret_cast<uint32_t*>(m.g + kChunkRows);
    m.offn = m.beg + kChunkRows;
  }
  __device__ __forceinline__ uint8_t* stage_ptr(int tile) const { return smem + L.main_bytes + (size_t)(tile % kStages) * L.stage_bytes; }

  // gather -> rowptr for every row of the chunk: thread t takes the rows base + t*rpl .. +rpl-1, so that one warp covers
  // exactly one tile and the offsets inside the tile are a warp scan
  __device__ __forceinline__ void build_table() {
    const int rows = ntiles * p.TR;
    for (int base = 0; base < rows; base += kThreads * rpl) {
      int64_t g[2] = {-1, -1};
      int row[2];
#pragma unroll
      for (int k = 0; k < 2; ++k) {
        row[k] = base + (int)threadIdx.x * rpl + k;
        if (k < rpl && row[k] < rows) {
          const int64_t r = r_begin + row[k];
          if (r < r_end) g[k] = p.gather ? p.gather[r] : r;
        }
      }
      int64_t beg[2] = {0, 0};
      int n[2] = {0, 0};
#pragma unroll
      for (int k = 0; k < 2; ++k)
        if (g[k] >= 0) {
          beg[k] = p.rowptr[g[k]];
          n[k] = min((int)(p.rowptr[g[k] + 1] - beg[k]), kMaxNnz);
        }
      const int np0 = (n[0] + pad_mask) & ~pad_mask, np1 = (n[1] + pad_mask) & ~pad_mask;
      const int s = np0 + np1;
      int v = s;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int u = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= o) v += u;
      }
      int off = v - s;
#pragma unroll
      for (int k = 0; k < 2; ++k)
        if (k < rpl && row[k] < rows) {
          m.g[row[k]] = (int)g[k];
          m.beg[row[k]] = (uint32_t)beg[k];
          m.offn[row[k]] = ((uint32_t)off << 8) | (uint32_t)n[k];
          off += k == 0 ? np0 : np1;
        }
    }
  }

  // ---- cp.async the entries / tail of a tile: thread t copies tile row t (a row has a handful of entries; a warp per row
  // with a lane per entry cost ten times the instructions); `extra` adds whatever else the kernel stages ----
  template <bool ROWID, typename Extra>
  __device__ __forceinline__ void stage(int tile, Extra& extra) {
    if (tile < ntiles) {
      uint8_t* sp = stage_ptr(tile);
      const int trow = threadIdx.x;
      if (trow < p.TR) {
        const int row = tile * p.TR + trow;
        const int g = m.g[row];
        if (g >= 0) {
          const int n = m.n(row), off = m.off(row);
          const int2* src = p.ent + m.beg[row];
          int2* dst = reinterpret_cast<int2*>(sp + L.ent_off) + off;
          uint8_t* rid = sp + L.rowid_off + off;
          for (int i = 0; i < n; ++i) {
            cp_async8(s32(dst + i), src + i);
            if (ROWID) rid[i] = (uint8_t)trow;
          }
          const int npad = (n + pad_mask) & ~pad_mask;
          for (int i = n; i < npad; ++i) dst[i] = make_int2(0, 0);   // zero-weight padding entries
          float* tl = reinterpret_cast<float*>(sp + L.tail_off) + trow * kMaxTail;
          const float* ts = p.tail + (int64_t)g * p.T;
          for (int t = 0; t < p.T; ++t) cp_async4(s32(tl + t), ts + t);
        }
      }
      extra(tile, sp);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");  // one group per tile, empty or not
  }

  template <bool ROWID, typename Extra, typename Process>
  __device__ __forceinline__ void run(Extra& extra, Process& process) {
    // a chunk longer than the table is walked in blocks of kChunkRows rows (two more global latencies per block)
    for (int64_t first = 0; first < chunk_end - chunk_begin; first += kChunkRows) {
      set_block(first);
      build_table();
      __syncthreads();   // table (and whatever the kernel wrote to shared memory before) visible
      stage<ROWID>(0, extra);
      stage<ROWID>(1, extra);
      for (int tile = 0; tile < ntiles; ++tile) {
        asm volatile("cp.async.wait_group 1;" ::: "memory");  // tile `tile` has landed (this thread's copies)
        __syncthreads();  // ... for everyone; tile - 1 is fully processed (its stage is free)
        stage<ROWID>(tile + 2, extra);
        process(tile, stage_ptr(tile), m, tile * p.TR);
      }
      asm volatile("cp.async.wait_group 0;" ::: "memory");
      __syncthreads();   // everyone is done with the table and the ring before the next block rewrites them
    }
  }
};

// ---- forward -----------------------------------------------------------------------------------------------------------
template <int LPR>
__global__ void __launch_bounds__(kThreads, 1) bag_fwd_kernel(BagP p) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  constexpr int SW = 4 * LPR;
  constexpr int kSlots = kThreads / LPR;
  const Layout L = make_layout((size_t)p.F * SW, p.TR, p.ecap, SW, false);
  float4* Ws = reinterpret_cast<float4*>(smem_raw);
  const int h0 = blockIdx.x * SW;
  for (int i = threadIdx.x; i < p.F * LPR; i += kThreads) {   // asynchronous: lands behind the index-chain prologue
    const int j = i / LPR, c = i - j * LPR;
    cp_async16(s32(Ws + i), p.WT + (int64_t)j * p.H + h0 + 4 * c, 16);
  }
  asm volatile("cp.async.commit_group;" ::: "memory");   // the oldest group: complete before tile 0 is consumed
  Ring ring(p, smem_raw, L, 3);
  const int l = threadIdx.x % LPR, slot = threadIdx.x / LPR;
  const float4 b4 = p.bias ? ld_f4(p.bias + h0 + 4 * l) : make_float4(0.f, 0.f, 0.f, 0.f);
  const bool drop = p.dropout_p > 0.f;
  const float keep_scale = drop ? 1.f / (1.f - p.dropout_p) : 1.f;
  const uint64_t rng_base = p.offset + ((drop && p.st) ? p.st->rng_offset : 0ull);
  const float4* Wl = Ws + l;

  auto none = [](int, uint8_t*) {};
  auto process = [&](int tile, uint8_t* sp, const MetaView& m, int crow0) {
    const int2* ent_s = reinterpret_cast<const int2*>(sp + L.ent_off);
    const float* tail_s = reinterpret_cast<const float*>(sp + L.tail_off);
    for (int row = slot; row < p.TR; row += kSlots) {
      const int64_t r = ring.r_begin + crow0 + row;
      if (r >= ring.r_end) break;
      const int n = (m.n(crow0 + row) + 3) & ~3;   // the stage pads every row to a multiple of four entries (zero weights)
      const int2* e = ent_s + m.off(crow0 + row);
      float4 acc = b4;
      for (int i = 0; i < n; i += 4) {  // entries are broadcast loads; four weight rows in flight
        const int2 e0 = e[i], e1 = e[i + 1], e2 = e[i + 2], e3 = e[i + 3];
        const float4 w0 = Wl[e0.x * LPR], w1 = Wl[e1.x * LPR], w2 = Wl[e2.x * LPR], w3 = Wl[e3.x * LPR];
        const float x0 = __int_as_float(e0.y), x1 = __int_as_float(e1.y), x2 = __int_as_float(e2.y), x3 = __int_as_float(e3.y);
        acc.x = fmaf(x0, w0.x, acc.x); acc.y = fmaf(x0, w0.y, acc.y); acc.z = fmaf(x0, w0.z, acc.z); acc.w = fmaf(x0, w0.w, acc.w);
        acc.x = fmaf(x1, w1.x, acc.x); acc.y = fmaf(x1, w1.y, acc.y); acc.z = fmaf(x1, w1.z, acc.z); acc.w = fmaf(x1, w1.w, acc.w);
        acc.x = fmaf(x2, w2.x, acc.x); acc.y = fmaf(x2, w2.y, acc.y); acc.z = fmaf(x2, w2.z, acc.z); acc.w = fmaf(x2, w2.w, acc.w);
        acc.x = fmaf(x3, w3.x, acc.x); acc.y = fmaf(x3, w3.y, acc.y); acc.z = fmaf(x3, w3.z, acc.z); acc.w = fmaf(x3, w3.w, acc.w);
      }
      for (int t = 0; t < p.T; ++t) {
        const float x0 = tail_s[row * kMaxTail + t];
        const float4 w0 = Wl[(p.tail_start + t) * LPR];
        acc.x = fmaf(x0, w0.x, acc.x); acc.y = fmaf(x0, w0.y, acc.y); acc.z = fmaf(x0, w0.z, acc.z); acc.w = fmaf(x0, w0.w, acc.w);
      }
      if (p.relu) {
        acc.x = fmaxf(acc.x, 0.f); acc.y = fmaxf(acc.y, 0.f); acc.z = fmaxf(acc.z, 0.f); acc.w = fmaxf(acc.w, 0.f);
      }
      if (drop) {
        // same element numbering as the GEMM epilogue (gemm_tc.cu): element e = r * H + h, keep bits of Philox(seed, e >> 2)
        const uint64_t e0 = rng_base + (uint64_t)r * (uint64_t)p.H + (uint64_t)(h0 + 4 * l);
        if ((e0 & 3) == 0) {  // the usual case (offsets and H are multiples of 4): one Philox block covers the four elements
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
  };
  ring.run<false>(none, process);
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
template <int SW>
__global__ void __launch_bounds__(kThreads, 1) bag_wgrad_kernel(BagP p) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  constexpr int CPL = SW / 32;  // accumulator columns per lane
  const int Fs = p.tail_start;
  const Layout L = make_layout((size_t)Fs * SW, p.TR, p.ecap, SW, true);
  float* acc = reinterpret_cast<float*>(smem_raw);  // [Fs][SW]
  for (int i = threadIdx.x; i < Fs * SW / 4; i += kThreads) reinterpret_cast<float4*>(acc)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  Ring ring(p, smem_raw, L, 0);
  const int lane = ring.lane, warp = ring.warp;
  const int h0 = blockIdx.x * SW;

  float tacc[kMaxTail + 1][CPL];  // dense tail columns + the bias gradient, over the rows this warp owns
#pragma unroll
  for (int t = 0; t <= kMaxTail; ++t)
#pragma unroll
    for (int c = 0; c < CPL; ++c) tacc[t][c] = 0.f;

  auto extra = [&](int tile, uint8_t* sp) {  // the dh slice of the tile: TR rows x SW floats in 16-byte pieces
    float* dh_s = reinterpret_cast<float*>(sp + L.dh_off);
    const int64_t r0 = ring.r_begin + (int64_t)tile * p.TR;
    for (int i = threadIdx.x; i < p.TR * (SW / 4); i += kThreads) {
      const int row = i / (SW / 4), c = i - row * (SW / 4);
      const int64_t r = r0 + row;
      const bool ok = r < ring.r_end;
      cp_async16(s32(dh_s + (size_t)row * SW + 4 * c), ok ? (const void*)(p.dh + r * p.lddh + h0 + 4 * c) : (const void*)p.dh, ok ? 16 : 0);
    }
  };
  auto process = [&](int tile, uint8_t* sp, const MetaView& m, int crow0) {
    const int2* ent_s = reinterpret_cast<const int2*>(sp + L.ent_off);
    const float* tail_s = reinterpret_cast<const float*>(sp + L.tail_off);
    const float* dh_s = reinterpret_cast<const float*>(sp + L.dh_off);
    const uint8_t* rowid_s = sp + L.rowid_off;
    const int E = m.off(crow0 + p.TR - 1) + m.n(crow0 + p.TR - 1);
    // sparse columns: flat scan, 32 entries at a time; this warp applies the entries of the columns it owns, in order
    for (int base = 0; base < E; base += 32) {
      int2 e = make_int2(-1, 0);
      int row = 0;
      if (base + lane < E) {
        e = ent_s[base + lane];
        row = rowid_s[base + lane];
      }
      unsigned mine = __ballot_sync(0xffffffffu, e.x >= 0 && (e.x & (kWarps - 1)) == warp);
      while (mine) {
        const int src = __ffs(mine) - 1;
        mine &= mine - 1;
        const int j = __shfl_sync(0xffffffffu, e.x, src);
        const float x = __int_as_float(__shfl_sync(0xffffffffu, e.y, src));
        const int rr = __shfl_sync(0xffffffffu, row, src);
#pragma unroll
        for (int c = 0; c < CPL; ++c) {
          float* a = acc + (size_t)j * SW + lane + 32 * c;
          *a = fmaf(x, dh_s[rr * SW + lane + 32 * c], *a);
        }
      }
    }
    // dense tail + bias gradient: warp w takes the tile rows w, w + 16, ...
    const int rows_here = (int)min((int64_t)p.TR, ring.r_end - (ring.r_begin + crow0));
    for (int row = warp; row < rows_here; row += kWarps) {
      float d[CPL];
#pragma unroll
      for (int c = 0; c < CPL; ++c) {
        d[c] = dh_s[row * SW + lane + 32 * c];
        tacc[kMaxTail][c] += d[c];
      }
#pragma unroll
      for (int t = 0; t < kMaxTail; ++t) {
        if (t < p.T) {
          const float x = tail_s[row * kMaxTail + t];
#pragma unroll
          for (int c = 0; c < CPL; ++c) tacc[t][c] = fmaf(x, d[c], tacc[t][c]);
        }
      }
    }
  };
  ring.run<true>(extra, process);  // the first barrier inside run() also covers the zeroed accumulator
  __syncthreads();
  float* part = p.partial + (size_t)blockIdx.y * (size_t)(p.F + 1) * p.H;
  for (int i = threadIdx.x; i < Fs * (SW / 4); i += kThreads) {
    const int j = i / (SW / 4), c = i - j * (SW / 4);
    st_f4(part + (size_t)j * p.H + h0 + 4 * c, reinterpret_cast<float4*>(acc)[i]);
  }
  // tail columns and bias: combine the warps' partial sums in warp order
  float* red = reinterpret_cast<float*>(smem_raw + L.main_bytes);  // [kWarps][kMaxTail + 1][SW]  (the ring is idle now)
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

// ---- host-side configuration ----------------------------------------------------------------------------------------
struct Config {
  int sw, TR, ecap;
  size_t smem;
};
// slice width and tile rows such that main region + 3 stages + meta fit in shared memory
static bool configure(int64_t H, int64_t F, int64_t tail_start, int64_t max_nnz, bool wgrad, Config* out) {
  if (max_nnz < 1) max_nnz = 1;
  if (max_nnz > kMaxNnz) return false;
  for (int sw : {64, 32}) {
    if (H % sw) continue;
    for (int TR : {64, 32}) {
      const int ecap = (int)(TR * (wgrad ? max_nnz : (max_nnz + 3) / 4 * 4));
      const Layout L = make_layout((size_t)(wgrad ? tail_start : F) * sw, TR, ecap, sw, wgrad);
      const size_t red = (size_t)kWarps * (kMaxTail + 1) * sw * 4;
      if (L.total <= kSmemLimit && (!wgrad || red <= kStages * L.stage_bytes)) {
        *out = Config{sw, TR, ecap, L.total};
        return true;
      }
    }
  }
  return false;
}
// rows per CTA: one chunk per SM - 153 CTAs on 148 SMs would be two waves (the kernel walks a chunk longer than its meta
// table in blocks)
static int64_t rows_per_chunk_for(int64_t R, int slices, int TR) {
  int64_t c = num_sms() / slices;
  if (c < 1) c = 1;
  int64_t rows = align_up(ceil_div(R, c), TR);
  if (rows < TR) rows = TR;
  return rows;
}

}  // namespace bag
}  // namespace ttam

using namespace ttam;
using namespace ttam::bag;

extern "C" int ttam_bag_supported(int64_t H, int64_t F, int64_t T, int64_t max_nnz) {
  Config a, b;
  return T >= 0 && T <= kMaxTail && T <= F && configure(H, F, F - T, max_nnz, false, &a) && configure(H, F, F - T, max_nnz, true, &b);
}

extern "C" int64_t ttam_bag_linear_workspace_bytes(int64_t R, int64_t H, int64_t F) {
  // forward: W^T; weight gradient: one [F+1][H] partial per row chunk (at most one chunk per SM)
  const int64_t fwd = align_up(F * H * 4, 256);
  // chunk count of the configuration with the most chunks (slices = H / 64 or H / 32, TR = 32)
  int64_t chunks = 1;
  for (int sw : {64, 32})
    if (H % sw == 0) {
      const int64_t c = ceil_div(R > 0 ? R : 1, rows_per_chunk_for(R > 0 ? R : 1, (int)(H / sw), 32));
      if (c > chunks) chunks = c;
    }
  const int64_t wg = chunks * (F + 1) * H * 4;
  return (fwd > wg ? fwd : wg) + 256;
}

static void fill_common(BagP& p, const int64_t* rowptr, const void* entries, const float* tail, int64_t T, int64_t tail_start,
                        const int64_t* gather, int64_t R, int64_t H, int64_t F, const Config& cfg) {
  p.rowptr = rowptr; p.ent = (const int2*)entries; p.tail = tail; p.T = (int)T; p.tail_start = (int)tail_start; p.gather = gather;
  p.R = R; p.F = (int)F; p.H = (int)H; p.TR = cfg.TR; p.ecap = cfg.ecap;
}

extern "C" int ttam_bag_linear_fwd(const int64_t* rowptr, const void* entries, const float* tail, int64_t T, int64_t tail_start,
                                   int64_t max_nnz, const int64_t* gather, int64_t R, const float* w, int64_t ldw, const float* bias,
                                   float* y, int64_t ldy, int64_t H, int64_t F, int act, float dropout_p, uint64_t seed,
                                   uint64_t offset, const void* state_dev, int round_tf32_out, void* workspace,
                                   int64_t workspace_bytes, void* stream) {
  if (R == 0) return TTAM_OK;
  TTAM_CHECK_ARG(rowptr && entries && w && y && workspace, "bag_linear_fwd: null pointer");
  TTAM_CHECK_ARG(act == TTAM_ACT_NONE || act == TTAM_ACT_RELU, "bag_linear_fwd: fuses ReLU only");
  TTAM_CHECK_ARG(T >= 0 && T <= kMaxTail && tail_start + T == F && (T == 0 || tail), "bag_linear_fwd: bad dense tail");
  Config cfg;
  TTAM_CHECK_ARG(configure(H, F, tail_start, max_nnz, false, &cfg),
                 "bag_linear_fwd: unsupported shape H=%lld F=%lld max_nnz=%lld (H %% 32 == 0, max_nnz <= 64, W slice + ring <= 226 kB)",
                 (long long)H, (long long)F, (long long)max_nnz);
  TTAM_CHECK_ARG(ldy % 4 == 0 && ((uintptr_t)y & 15) == 0 && (!bias || ((uintptr_t)bias & 15) == 0), "bag_linear_fwd: output / bias must be 16-byte aligned");
  if (workspace_bytes < align_up(F * H * 4, 256)) {  // the forward only needs room for W^T
    set_error("bag_linear_fwd: workspace too small");
    return TTAM_EWORKSPACE;
  }
  cudaStream_t s = (cudaStream_t)stream;
  float* WT = (float*)workspace;
  transpose_w_kernel<<<dim3((unsigned)ceil_div(F, 32), (unsigned)ceil_div(H, 32)), dim3(32, 8), 0, s>>>(w, ldw, (int)H, (int)F, WT);
  TTAM_LAUNCH_CHECK();
  BagP p{};
  fill_common(p, rowptr, entries, tail, T, tail_start, gather, R, H, F, cfg);
  p.WT = WT; p.bias = bias; p.y = y; p.ldy = ldy; p.relu = act == TTAM_ACT_RELU;
  p.dropout_p = dropout_p; p.seed = seed; p.offset = offset; p.st = (const ttam_step_state*)state_dev; p.round_out = round_tf32_out;
  const int slices = (int)(H / cfg.sw);
  p.rows_per_chunk = rows_per_chunk_for(R, slices, cfg.TR);
  dim3 grid((unsigned)slices, (unsigned)ceil_div(R, p.rows_per_chunk));
  if (cfg.sw == 64) {
    static bool done = false;
    if (!done) { TTAM_CUDA(cudaFuncSetAttribute(bag_fwd_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemLimit)); done = true; }
    bag_fwd_kernel<16><<<grid, kThreads, cfg.smem, s>>>(p);
  } else {
    static bool done = false;
    if (!done) { TTAM_CUDA(cudaFuncSetAttribute(bag_fwd_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemLimit)); done = true; }
    bag_fwd_kernel<8><<<grid, kThreads, cfg.smem, s>>>(p);
  }
  TTAM_LAUNCH_CHECK();
  return TTAM_OK;
}

extern "C" int ttam_bag_linear_wgrad(const int64_t* rowptr, const void* entries, const float* tail, int64_t T, int64_t tail_start,
                                     int64_t max_nnz, const int64_t* gather, int64_t R, const float* dh, int64_t lddh, float* dw,
                                     int64_t lddw, float* db, int64_t H, int64_t F, int accumulate, void* workspace,
                                     int64_t workspace_bytes, void* stream) {
  TTAM_CHECK_ARG(rowptr && entries && (dh || R == 0) && dw && workspace, "bag_linear_wgrad: null pointer");
  TTAM_CHECK_ARG(T >= 0 && T <= kMaxTail && tail_start + T == F && (T == 0 || tail), "bag_linear_wgrad: bad dense tail");
  Config cfg;
  TTAM_CHECK_ARG(configure(H, F, tail_start, max_nnz, true, &cfg), "bag_linear_wgrad: unsupported shape H=%lld F=%lld max_nnz=%lld",
                 (long long)H, (long long)F, (long long)max_nnz);
  TTAM_CHECK_ARG(lddh % 4 == 0 && ((uintptr_t)dh & 15) == 0, "bag_linear_wgrad: dh must be 16-byte aligned");
  if (workspace_bytes < ttam_bag_linear_workspace_bytes(R, H, F)) {
    set_error("bag_linear_wgrad: workspace too small");
    return TTAM_EWORKSPACE;
  }
  cudaStream_t s = (cudaStream_t)stream;
  const int slices = (int)(H / cfg.sw);
  if (R > 0) {
    BagP p{};
    fill_common(p, rowptr, entries, tail, T, tail_start, gather, R, H, F, cfg);
    p.dh = dh; p.lddh = lddh; p.partial = (float*)workspace;
    p.rows_per_chunk = rows_per_chunk_for(R, slices, cfg.TR);
    dim3 grid((unsigned)slices, (unsigned)ceil_div(R, p.rows_per_chunk));
    const int64_t real_chunks = grid.y;
    if (cfg.sw == 64) {
      static bool done = false;
      if (!done) { TTAM_CUDA(cudaFuncSetAttribute(bag_wgrad_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemLimit)); done = true; }
      bag_wgrad_kernel<64><<<grid, kThreads, cfg.smem, s>>>(p);
    } else {
      static bool done = false;
      if (!done) { TTAM_CUDA(cudaFuncSetAttribute(bag_wgrad_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemLimit)); done = true; }
      bag_wgrad_kernel<32><<<grid, kThreads, cfg.smem, s>>>(p);
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
