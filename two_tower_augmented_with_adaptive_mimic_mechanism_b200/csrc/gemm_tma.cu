// TMA-fed, persistent, warp-specialised TF32 GEMM (tcgen05.mma kind::tf32, fp32 accumulate in TMEM) for the K-major x
// K-major products of the tower chain: the forward GEMMs  y = x . w^T  of nn.Linear (reference encoders.py:121-144,157-162)
// on activation matrices that are NOT row-gathered, and their data gradients  dx = dy . w  taken against a transposed copy of
// the weight (ttam_prepare_weights), which makes them the same product.
//
//   C[M, N] = A[M, K] . B[N, K]^T      A: activations, rows lda floats apart;  B: weight (or transposed weight), pre-rounded to TF32
//
// Why a second GEMM kernel: gemm_tc.cu stages its operands with 8 producer warps (cp.async + an in-place rounding pass +
// a per-chunk mbarrier hand-shake) because it also has to gather rows and to lay out MN-major operands; for plain K-major
// operands that machinery is the bottleneck (ncu: 1.3 % tensor pipe, 12 % DRAM, latency-bound: profiles/r1_gemm_*).  Here
//   warp 0      one thread issues cp.async.bulk.tensor (TMA, 128-byte swizzle) for A and B chunks of 32 reduction elements into
//               a ring of shared-memory stages - no registers, no per-thread copies, the swizzled UMMA layout comes for free
//   warp 1      one thread issues 4 tcgen05.mma (M = 128, N = bn <= 256, K = 8) per chunk; accumulators are DOUBLE-BUFFERED in
//               TMEM, so the tensor core starts tile t+1 while tile t is drained
//   warps 2-5   epilogue: tcgen05.ld (thread = output row), bias / ReLU / ReLU-mask / scale / accumulate / Philox dropout /
//               optional TF32 rounding of the output, 16-byte stores
//   warps 6-9   only when A is not pre-rounded: round the landed A chunk to TF32 (nearest) in place before the MMA reads it -
//               the tensor core would truncate (same arithmetic as gemm_tc.cu: results are bit-identical)
// CTAs are persistent (one per SM) and walk the (M tile, N tile) units round-robin.
#include "gemm_tc.cuh"
#include "sm100.cuh"
#include <cstdlib>

namespace ttam {
namespace tma {

using namespace ttam::sm100;

constexpr int kBM = 128;
constexpr int kKC = 32;                   // reduction elements per stage: one 128-byte swizzled row of fp32
constexpr uint32_t kABytes = kBM * 128;   // 16 KB
constexpr int kEpiWarps = 8, kRoundWarps = 4;   // forward / dgrad kernel: one set of four epilogue warps per accumulator buffer
constexpr int kThreads = 32 * (2 + kEpiWarps + kRoundWarps);
constexpr int kEpiWarpsW = 4;                   // weight-gradient kernel
constexpr int kThreadsW = 32 * (2 + kEpiWarpsW + kRoundWarps);
constexpr int kMaxStages = 8;

struct TmaP {
  float* C;
  int64_t ldc;
  int M, N, K;
  int bn, acc_stride, n_tiles_n, units, nchunks, stages;
  int roundA, round_out, vecC, vecAux, vecBias;
  const float* bias;
  int relu;
  float dropout_p;
  uint64_t seed, offset;
  const ttam_step_state* st;
  const float* aux;
  int64_t ldaux;
  int mask_mode;
  float scale;
  int accumulate;
};

__device__ __forceinline__ uint32_t to_tf32(float x) {
  const uint32_t u = __float_as_uint(x);
  return ((u & 0x7F800000u) != 0x7F800000u) ? u + 0x1000u : u;
}
__device__ __forceinline__ float round_out_tf32(float x) {
  const uint32_t u = __float_as_uint(x);
  return __uint_as_float(((u & 0x7F800000u) != 0x7F800000u) ? ((u + 0x1000u) & 0xFFFFE000u) : u);
}

// One 32 x 32 chunk of a tile that lies fully inside C, from the warp's staging tile to global memory: lane = 4 columns of
// the rows sub_row, sub_row + 4, ... (8 lanes cover the 128 contiguous bytes of a row, 4 rows per instruction).  The
// variant is chosen once per launch, so that the row loop carries no tests: with them the chunk took ~1300 instructions,
// and four epilogue warps - one per scheduler, nothing to switch to - were what bounded the kernel (ncu: issue-bound
// epilogue, DRAM 25 %, the same 26 us for K = 96 and K = 192).
template <bool RELU, bool MASK, bool ACC, bool SCALE, bool ROUND>
__device__ __forceinline__ void epi_chunk_full(const float* __restrict__ stg_lane, float* __restrict__ c_lane, int64_t ldc4,
                                               const float* __restrict__ aux_lane, int64_t ldaux4, const float4 b4, const float scale) {
  float4 auxv[8], oldv[8];
  if (MASK) {
#pragma unroll
    for (int i = 0; i < 8; ++i) auxv[i] = ld_f4(aux_lane + i * ldaux4);
  }
  if (ACC) {
#pragma unroll
    for (int i = 0; i < 8; ++i) oldv[i] = ld_f4(c_lane + i * ldc4);
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const float4 a = *reinterpret_cast<const float4*>(stg_lane + i * (4 * 36));
    float4 o = make_float4(a.x + b4.x, a.y + b4.y, a.z + b4.z, a.w + b4.w);
    if (RELU) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
    if (MASK) {
      o.x = auxv[i].x > 0.f ? o.x : 0.f; o.y = auxv[i].y > 0.f ? o.y : 0.f;
      o.z = auxv[i].z > 0.f ? o.z : 0.f; o.w = auxv[i].w > 0.f ? o.w : 0.f;
    }
    if (SCALE) { o.x *= scale; o.y *= scale; o.z *= scale; o.w *= scale; }
    if (ACC) { o.x += oldv[i].x; o.y += oldv[i].y; o.z += oldv[i].z; o.w += oldv[i].w; }
    if (ROUND) { o.x = round_out_tf32(o.x); o.y = round_out_tf32(o.y); o.z = round_out_tf32(o.z); o.w = round_out_tf32(o.w); }
    st_f4(c_lane + i * ldc4, o);
  }
}

#define TTAM_EPI_CASE(m) \
  case m: epi_chunk_full<((m) & 1) != 0, ((m) & 2) != 0, ((m) & 4) != 0, ((m) & 8) != 0, ((m) & 16) != 0>(stg_lane, c_lane, ldc4, aux_lane, ldaux4, b4, scale); break;
__device__ __forceinline__ void epi_chunk_dispatch(int mode, const float* __restrict__ stg_lane, float* __restrict__ c_lane, int64_t ldc4,
                                                   const float* __restrict__ aux_lane, int64_t ldaux4, const float4 b4, const float scale) {
  switch (mode) {
    TTAM_EPI_CASE(0) TTAM_EPI_CASE(1) TTAM_EPI_CASE(2) TTAM_EPI_CASE(3) TTAM_EPI_CASE(4) TTAM_EPI_CASE(5) TTAM_EPI_CASE(6) TTAM_EPI_CASE(7)
    TTAM_EPI_CASE(8) TTAM_EPI_CASE(9) TTAM_EPI_CASE(10) TTAM_EPI_CASE(11) TTAM_EPI_CASE(12) TTAM_EPI_CASE(13) TTAM_EPI_CASE(14) TTAM_EPI_CASE(15)
    TTAM_EPI_CASE(16) TTAM_EPI_CASE(17) TTAM_EPI_CASE(18) TTAM_EPI_CASE(19) TTAM_EPI_CASE(20) TTAM_EPI_CASE(21) TTAM_EPI_CASE(22) TTAM_EPI_CASE(23)
    TTAM_EPI_CASE(24) TTAM_EPI_CASE(25) TTAM_EPI_CASE(26) TTAM_EPI_CASE(27) TTAM_EPI_CASE(28) TTAM_EPI_CASE(29) TTAM_EPI_CASE(30) TTAM_EPI_CASE(31)
  }
}
#undef TTAM_EPI_CASE

__global__ void __launch_bounds__(kThreads, 1)
gemm_tf32_tma_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, TmaP p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const uint32_t stage_bytes = kABytes + (uint32_t)p.bn * 128u;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (uint32_t)p.stages * stage_bytes);
  uint64_t* full = bars;                          // [stages]  TMA bytes landed
  uint64_t* ready = bars + kMaxStages;            // [stages]  A chunk rounded (only when roundA)
  uint64_t* empty = bars + 2 * kMaxStages;        // [stages]  MMAs that read the stage have completed
  uint64_t* acc_full = bars + 3 * kMaxStages;     // [2]
  uint64_t* acc_empty = bars + 3 * kMaxStages + 2;  // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3 * kMaxStages + 4);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t tmem_cols = 2 * p.acc_stride <= 32 ? 32u : 2 * p.acc_stride <= 64 ? 64u : 2 * p.acc_stride <= 128 ? 128u
                             : 2 * p.acc_stride <= 256 ? 256u : 512u;
  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int i = 0; i < p.stages; ++i) {
      mbar_init(full + i, 1);
      mbar_init(ready + i, kRoundWarps);
      mbar_init(empty + i, 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(acc_full + i, 1);
      mbar_init(acc_empty + i, 4);
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      uint32_t it = 0;
      for (int u = blockIdx.x; u < p.units; u += gridDim.x) {
        const int mt = u / p.n_tiles_n, nt = u - mt * p.n_tiles_n;
        for (int c = 0; c < p.nchunks; ++c, ++it) {
          const int stage = it % p.stages;
          mbar_wait_relaxed(empty + stage, ((it / p.stages) & 1) ^ 1, 64);
          uint8_t* sA = smem + (uint32_t)stage * stage_bytes;
          mbar_expect_tx(full + stage, stage_bytes);
          tma_load_2d(sA, &tmA, full + stage, c * kKC, mt * kBM);
          tma_load_2d(sA + kABytes, &tmB, full + stage, c * kKC, nt * p.bn);
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      const uint32_t idesc = make_idesc(2 /*tf32*/, kBM, p.bn, 0, 0);
      uint32_t it = 0, un = 0;
      for (int u = blockIdx.x; u < p.units; u += gridDim.x, ++un) {
        const int buf = un & 1;
        mbar_wait(acc_empty + buf, ((un >> 1) & 1) ^ 1);   // the epilogue has drained this accumulator
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(buf * p.acc_stride);
        for (int c = 0; c < p.nchunks; ++c, ++it) {
          const int stage = it % p.stages;
          mbar_wait((p.roundA ? ready : full) + stage, (it / p.stages) & 1);
          tc_fence_after();
          const uint32_t sA = smem_u32(smem + (uint32_t)stage * stage_bytes);
          const uint32_t sB = sA + kABytes;
#pragma unroll
          for (int ks = 0; ks < kKC / 8; ++ks)
            umma_tf32(d_tmem, make_kmajor_desc<128>(sA + ks * 32), make_kmajor_desc<128>(sB + ks * 32), idesc, (c | ks) != 0 ? 1u : 0u);
          umma_commit(empty + stage);
        }
        umma_commit(acc_full + buf);
      }
    }
  } else if (warp < 2 + kEpiWarps) {
    // ===================== epilogue =====================
    // tcgen05.ld hands every thread one accumulator ROW (32 columns of it per load).  Stored like that, a warp's store would
    // touch 32 rows x 16 bytes - half-used sectors on 32 different lines - and the mask / accumulate operands would be read
    // the same way (measured: the kernel ran at the speed of this epilogue, 2x slower for 2x the columns).  So each warp
    // transposes its 32 x 32 block through a private shared-memory tile (row pitch 36 floats: conflict-free both ways) and
    // does all global traffic with 8 lanes per row: 128 contiguous bytes per row, 4 rows per instruction.
    const int quad = warp & 3;   // the TMEM lanes this warp may read: 32*quad .. +31
    const int set = (warp - 2) >> 2;   // which accumulator buffer this warp drains (buffer = unit count & 1)
    float* stg = reinterpret_cast<float*>(smem + (uint32_t)p.stages * stage_bytes + 512) + (warp - 2) * (32 * 36);
    const bool drop = p.dropout_p > 0.f;
    const float keep_scale = drop ? 1.f / (1.f - p.dropout_p) : 1.f;
    const uint64_t rng_base = p.offset + ((drop && p.st) ? p.st->rng_offset : 0ull);
    const int sub_row = lane >> 3, c4 = (lane & 7) * 4;
    const bool masked = p.mask_mode == 1;
    const bool full_ok = p.vecC && !drop && (!masked || p.vecAux) && (!p.bias || p.vecBias);
    const int mode = (p.relu ? 1 : 0) | (masked ? 2 : 0) | (p.accumulate ? 4 : 0) | (p.scale != 1.f ? 8 : 0) | (p.round_out ? 16 : 0);
    const float* stg_lane = stg + sub_row * 36 + c4;
    const int64_t ldc4 = 4 * p.ldc, ldaux4 = 4 * p.ldaux;
    const float scale = p.scale;
    uint32_t un = 0;
    for (int u = blockIdx.x; u < p.units; u += gridDim.x, ++un) {
      if ((int)(un & 1) != set) continue;
      const int mt = u / p.n_tiles_n, nt = u - mt * p.n_tiles_n;
      const int buf = un & 1;
      const int m_base = mt * kBM + quad * 32, n0 = nt * p.bn;
      const int n_end = min(p.N, n0 + p.bn);
      mbar_wait(acc_full + buf, (un >> 1) & 1);
      tc_fence_after();
      for (int col = 0; col < p.bn; col += 32) {
        const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(buf * p.acc_stride + col);
        uint32_t r[32];
        if (col + 32 <= p.bn) {
          tmem_ld_32x32(taddr, r);
        } else {   // bn is a multiple of 16: a last half chunk
          uint32_t h[16];
          tmem_ld_32x16(taddr, h);
#pragma unroll
          for (int i = 0; i < 16; ++i) { r[i] = h[i]; r[16 + i] = 0u; }
        }
#pragma unroll
        for (int i = 0; i < 32; i += 4) *reinterpret_cast<uint4*>(stg + lane * 36 + i) = make_uint4(r[i], r[i + 1], r[i + 2], r[i + 3]);
        __syncwarp();
        const int n = n0 + col + c4;            // this lane's 4 columns in every row of the chunk
        if (full_ok && m_base + 32 <= p.M && n0 + col + 32 <= n_end) {   // the chunk lies inside C: no tests in the row loop
          const float4 b4 = p.bias ? ld_f4(p.bias + n) : make_float4(0.f, 0.f, 0.f, 0.f);
          epi_chunk_dispatch(mode, stg_lane, p.C + (int64_t)(m_base + sub_row) * p.ldc + n, ldc4,
                             masked ? p.aux + (int64_t)(m_base + sub_row) * p.ldaux + n : nullptr, ldaux4, b4, scale);
          __syncwarp();  // the tile is rewritten by the next chunk
          continue;
        }
        const bool in4 = n + 3 < p.N;
        float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
        if (p.bias && n < p.N) {
          if (p.vecBias && in4) b4 = ld_f4(p.bias + n);
          else {
            b4.x = p.bias[n];
            if (n + 1 < p.N) b4.y = p.bias[n + 1];
            if (n + 2 < p.N) b4.z = p.bias[n + 2];
            if (n + 3 < p.N) b4.w = p.bias[n + 3];
          }
        }
        const bool lane_ok = n < p.N && col + c4 < p.bn;
        const bool fast = in4 && p.vecC && (p.mask_mode != 1 || p.vecAux) && !drop;
        if (fast) {
          // all global operands of the 8 rows this lane touches are requested before the first one is used
          float4 auxv[8], oldv[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int m = m_base + i * 4 + sub_row;
            const bool ok = lane_ok && m < p.M;
            auxv[i] = (ok && p.mask_mode == 1) ? ld_f4(p.aux + (int64_t)m * p.ldaux + n) : make_float4(1.f, 1.f, 1.f, 1.f);
            oldv[i] = (ok && p.accumulate) ? ld_f4(p.C + (int64_t)m * p.ldc + n) : make_float4(0.f, 0.f, 0.f, 0.f);
          }
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int row = i * 4 + sub_row;
            const int m = m_base + row;
            if (lane_ok && m < p.M) {
              const float4 a = *reinterpret_cast<const float4*>(stg + row * 36 + c4);
              float4 o = make_float4(a.x + b4.x, a.y + b4.y, a.z + b4.z, a.w + b4.w);
              if (p.relu) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
              o.x = auxv[i].x > 0.f ? o.x : 0.f; o.y = auxv[i].y > 0.f ? o.y : 0.f;
              o.z = auxv[i].z > 0.f ? o.z : 0.f; o.w = auxv[i].w > 0.f ? o.w : 0.f;
              if (p.scale != 1.f) { o.x *= p.scale; o.y *= p.scale; o.z *= p.scale; o.w *= p.scale; }
              o.x += oldv[i].x; o.y += oldv[i].y; o.z += oldv[i].z; o.w += oldv[i].w;
              if (p.round_out) {
                o.x = round_out_tf32(o.x); o.y = round_out_tf32(o.y); o.z = round_out_tf32(o.z); o.w = round_out_tf32(o.w);
              }
              st_f4(p.C + (int64_t)m * p.ldc + n, o);
            }
          }
        } else {
#pragma unroll 1
        for (int i = 0; i < 8; ++i) {
          const int row = i * 4 + sub_row;
          const int m = m_base + row;
          if (m < p.M && lane_ok) {
            const float4 a = *reinterpret_cast<const float4*>(stg + row * 36 + c4);
            float v[4] = {a.x + b4.x, a.y + b4.y, a.z + b4.z, a.w + b4.w};
            if (p.relu) {
#pragma unroll
              for (int e = 0; e < 4; ++e) v[e] = fmaxf(v[e], 0.f);
            }
            if (drop) {
#pragma unroll
              for (int e = 0; e < 4; ++e)
                if (n + e < p.N) v[e] = dropout_keep(p.seed, rng_base + (uint64_t)m * (uint64_t)p.N + (uint64_t)(n + e), p.dropout_p) ? v[e] * keep_scale : 0.f;
            }
            if (p.mask_mode == 1) {
              const float* ax = p.aux + (int64_t)m * p.ldaux + n;
#pragma unroll
              for (int e = 0; e < 4; ++e)
                if (n + e < p.N) v[e] = ax[e] > 0.f ? v[e] : 0.f;
            }
            if (p.scale != 1.f) {
#pragma unroll
              for (int e = 0; e < 4; ++e) v[e] *= p.scale;
            }
            float* dst = p.C + (int64_t)m * p.ldc + n;
#pragma unroll
            for (int e = 0; e < 4; ++e)
              if (n + e < p.N) {
                const float o = p.accumulate ? dst[e] + v[e] : v[e];
                dst[e] = p.round_out ? round_out_tf32(o) : o;
              }
          }
        }
        }
        __syncwarp();  // the tile is rewritten by the next chunk
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(acc_empty + buf);
    }
  } else if (p.roundA) {
    // ===================== rounding warps: A chunk -> nearest TF32, in place =====================
    const int t = threadIdx.x - 32 * (2 + kEpiWarps);   // 0 .. 127
    uint32_t it = 0;
    for (int u = blockIdx.x; u < p.units; u += gridDim.x) {
      for (int c = 0; c < p.nchunks; ++c, ++it) {
        const int stage = it % p.stages;
        mbar_wait(full + stage, (it / p.stages) & 1);
        uint8_t* sA = smem + (uint32_t)stage * stage_bytes;
#pragma unroll
        for (int i = 0; i < (int)(kABytes / 16) / (32 * kRoundWarps); ++i) {
          uint4* q = reinterpret_cast<uint4*>(sA) + t + 32 * kRoundWarps * i;
          const float4 v = *reinterpret_cast<float4*>(q);
          *q = make_uint4(to_tf32(v.x), to_tf32(v.y), to_tf32(v.z), to_tf32(v.w));
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(ready + stage);
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, tmem_cols);
}

// ---------------------------------------------------------------------------------------------------------------------------
// Weight gradient  dW[N, K] = dy[R, N]^T . x[R, K]  (reduction over the rows r, split over blockIdx.y), both operands MN-major:
// the reduction index r is the ROW of both row-major matrices.  TMA lays the tiles out for the tensor core directly: a box of
// 32 columns x 32 rows with CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B is exactly one "MN-major, 128-byte swizzle, 32-byte atom" block of
// the UMMA layout (32 M/N elements x 32 reduction rows, rows 128 bytes apart, the 32-byte chunk index XOR-ed with the row index
// mod 4) - the layout gemm_tc.cu builds with 8 warps of cp.async.  C(m = feature k', n = n') is stored transposed,
// part[z][n'][k'], and summed over z by splitk_reduce_kernel like the partials of gemm_tc.cu.  The bias gradient (exact fp32
// column sums of dy, taken before the TF32 rounding) comes out of the same pass: the rounding warps accumulate it.
struct WgP {
  float* part;     // [splits][N][K]
  float* colsum;   // [splits][N] or null
  int R, N, K;
  int bn, m_tiles, rows_per_split, stages;
  int roundA, roundB;
};
constexpr uint32_t kAtomBytes = 32 * 128;   // one 32 x 32 fp32 block

__global__ void __launch_bounds__(kThreadsW, 1)
gemm_tf32_tma_wgrad_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmDy, WgP p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int nb_atoms = p.bn >> 5;
  const uint32_t stage_bytes = kABytes + (uint32_t)nb_atoms * kAtomBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (uint32_t)p.stages * stage_bytes);
  uint64_t* full = bars;
  uint64_t* ready = bars + kMaxStages;
  uint64_t* empty = bars + 2 * kMaxStages;
  uint64_t* acc_full = bars + 3 * kMaxStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3 * kMaxStages + 4);
  float* tail = reinterpret_cast<float*>(smem + (uint32_t)p.stages * stage_bytes + 512);   // epilogue tiles / colsum reduction

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int mt = blockIdx.x % p.m_tiles, nt = blockIdx.x / p.m_tiles;
  const int m0 = mt * kBM, n0 = nt * p.bn;
  const int r_lo = blockIdx.y * p.rows_per_split, r_hi = min(p.R, r_lo + p.rows_per_split);
  const int nchunks = (r_hi - r_lo + kKC - 1) / kKC;
  const bool want_colsum = p.colsum != nullptr && mt == 0;
  const bool need_round = p.roundA || p.roundB || want_colsum;
  const uint32_t tmem_cols = p.bn <= 32 ? 32u : p.bn <= 64 ? 64u : p.bn <= 128 ? 128u : 256u;
  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmX);
    tma_prefetch_desc(&tmDy);
    for (int i = 0; i < p.stages; ++i) {
      mbar_init(full + i, 1);
      mbar_init(ready + i, kRoundWarps);
      mbar_init(empty + i, 1);
    }
    mbar_init(acc_full, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      for (int c = 0; c < nchunks; ++c) {
        const int stage = c % p.stages;
        mbar_wait_relaxed(empty + stage, ((c / p.stages) & 1) ^ 1, 64);
        uint8_t* sA = smem + (uint32_t)stage * stage_bytes;
        mbar_expect_tx(full + stage, stage_bytes);
        const int r = r_lo + c * kKC;
        for (int a = 0; a < 4; ++a) tma_load_2d(sA + a * kAtomBytes, &tmX, full + stage, m0 + 32 * a, r);
        for (int b = 0; b < nb_atoms; ++b) tma_load_2d(sA + kABytes + b * kAtomBytes, &tmDy, full + stage, n0 + 32 * b, r);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = make_idesc(2 /*tf32*/, kBM, p.bn, 1, 1);
      for (int c = 0; c < nchunks; ++c) {
        const int stage = c % p.stages;
        mbar_wait((need_round ? ready : full) + stage, (c / p.stages) & 1);
        tc_fence_after();
        const uint32_t sA = smem_u32(smem + (uint32_t)stage * stage_bytes);
        const uint32_t sB = sA + kABytes;
#pragma unroll
        for (int ks = 0; ks < kKC / 8; ++ks)   // one MMA consumes 8 reduction rows: two 4-row groups (1024 B) of every block
          umma_tf32(tmem_base, make_mnmajor_desc_tf32(sA + ks * 1024, kAtomBytes, 512), make_mnmajor_desc_tf32(sB + ks * 1024, kAtomBytes, 512),
                    idesc, (c | ks) != 0 ? 1u : 0u);
        umma_commit(empty + stage);
      }
      umma_commit(acc_full);
    }
  } else if (warp < 2 + kEpiWarpsW) {
    // ===================== epilogue: transposed store  part[z][n][m]  through a [32][33] tile per warp =====================
    const int quad = warp & 3;
    float* stg = tail + (warp - 2) * (32 * 33);
    float* Cz = p.part + (int64_t)blockIdx.y * p.N * p.K;
    mbar_wait(acc_full, 0);
    tc_fence_after();
    for (int col = 0; col < p.bn; col += 32) {
      uint32_t r[32];
      tmem_ld_32x32(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)col, r);
#pragma unroll
      for (int i = 0; i < 32; ++i) stg[lane * 33 + i] = __uint_as_float(r[i]);
      __syncwarp();
      // lanes = 32 consecutive m: one 128-byte store per output row n; address and bounds hoisted out of the row loop (with
      // them inside, the loop body was ~70 instructions and the epilogue of a 128 x 96 tile took ~7 us of a ~20 us CTA)
      const int m = m0 + quad * 32 + lane;
      const int rows = min(32, p.N - (n0 + col));
      if (m < p.K) {
        float* dst = Cz + (int64_t)(n0 + col) * p.K + m;
        const float* srow = stg + lane * 33;
#pragma unroll 8
        for (int i = 0; i < 32; ++i)
          if (i < rows) dst[(int64_t)i * p.K] = srow[i];
      }
      __syncwarp();
    }
  } else {
    // ===================== rounding warps (+ exact column sums of dy) =====================
    const int t = threadIdx.x - 32 * (2 + kEpiWarpsW);   // 0 .. 127
    // B pieces of this thread: piece t + 128 i -> block i >> 1, reduction row (t >> 3) + 16 (i & 1), physical 16-byte chunk t & 7
    // -> logical chunk ((c >> 1) ^ (row & 3)) << 1 | (c & 1), the same for every i
    const int c16 = ((((t & 7) >> 1) ^ ((t >> 3) & 3)) << 1) | (t & 1);
    float4 csum[8];
#pragma unroll
    for (int b = 0; b < 8; ++b) csum[b] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (need_round) {
      for (int c = 0; c < nchunks; ++c) {
        const int stage = c % p.stages;
        mbar_wait(full + stage, (c / p.stages) & 1);
        uint8_t* sA = smem + (uint32_t)stage * stage_bytes;
        if (p.roundA) {
#pragma unroll
          for (int i = 0; i < (int)(kABytes / 16) / (32 * kRoundWarps); ++i) {
            uint4* q = reinterpret_cast<uint4*>(sA) + t + 32 * kRoundWarps * i;
            const float4 v = *reinterpret_cast<float4*>(q);
            *q = make_uint4(to_tf32(v.x), to_tf32(v.y), to_tf32(v.z), to_tf32(v.w));
          }
        }
        if (p.roundB || want_colsum) {
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            if (i < 2 * nb_atoms) {
              uint4* q = reinterpret_cast<uint4*>(sA + kABytes) + t + 128 * i;
              const float4 v = *reinterpret_cast<float4*>(q);
              if (want_colsum) { csum[i >> 1].x += v.x; csum[i >> 1].y += v.y; csum[i >> 1].z += v.z; csum[i >> 1].w += v.w; }
              if (p.roundB) *q = make_uint4(to_tf32(v.x), to_tf32(v.y), to_tf32(v.z), to_tf32(v.w));
            }
          }
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(ready + stage);
      }
    }
    if (want_colsum) {
      // combine the 16 threads that hold the same logical chunk, in thread order (deterministic)
      mbar_wait(acc_full, 0);   // every MMA has read the stages: the ring memory is free
      float4* red = reinterpret_cast<float4*>(smem);   // [128 threads][8 blocks]
#pragma unroll
      for (int b = 0; b < 8; ++b) red[t * 8 + b] = csum[b];
      asm volatile("bar.sync 1, %0;" ::"n"(32 * kRoundWarps) : "memory");
      for (int n = t; n < p.bn; n += 32 * kRoundWarps) {
        const int b = n >> 5, want = (n & 31) >> 2, comp = n & 3;
        float sacc = 0.f;
        for (int u = 0; u < 32 * kRoundWarps; ++u) {
          const int cu = ((((u & 7) >> 1) ^ ((u >> 3) & 3)) << 1) | (u & 1);
          if (cu == want) sacc += reinterpret_cast<const float*>(red + u * 8 + b)[comp];
        }
        if (n0 + n < p.N) p.colsum[(int64_t)blockIdx.y * p.N + n0 + n] = sacc;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, tmem_cols);
}

// ---------------------------------------------------------------------------------------------------------------------------
// Weight gradient of the BAG-form layer 1 on the tensor cores:  dW1[H, F] = dh[R, H]^T . X[idx][R, F]  with X given as CSR rows
// + a dense tail (csrc/bag.cu).  Same MMA, descriptors, epilogue and split-K reduce as gemm_tf32_tma_wgrad_kernel; what differs is
// where the A operand comes from and how the work is cut:
//   * instead of TMA boxes of a dense x, the four fill warps EXPAND the 32 CSR rows of a chunk into the shared-memory stage:
//     the handful of entries (and tail values) of a row are scattered straight into the "MN-major, 128-byte swizzle, 32-byte
//     atom" layout the UMMA descriptor expects (element (feature f, row r) of a 32 x 32 block at byte r*128 + c*16 + (f%4)*4
//     with the 16-byte chunk c = (((f/4)>>1 ^ (r&3)) << 1) | ((f/4)&1): what TMA produces with SWIZZLE_128B_ATOM_32B).  The
//     stages are zeroed once; when a stage comes round again each thread writes zeros back to exactly the words it scattered
//     there, so a chunk costs a few dozen shared-memory stores instead of a pass over the tile.
//   * a CTA owns ALL feature tiles (m_tiles x 128 features, one TMEM accumulator of bn columns each: m_tiles * bn <= 512) of one
//     slice of bn hidden columns and one split of the rows: every dh element is fetched and rounded by exactly one CTA, every
//     CSR row is read by H / bn CTAs.  One CTA per SM.
//   * the index chain idx -> rowptr of the split's rows is resolved once into shared memory; the entries of chunk c+1 are
//     fetched into registers while chunk c is scattered (4 threads per row: thread `sub` holds the entries sub, sub+4 and the
//     tail values sub, sub+4; longer rows take a slow loop).  Values are rounded to TF32 like a dense x.
// The SIMT kernel this replaces on the tensor-core path (bag_wgrad_kernel: per-entry read-modify-writes of a shared-memory
// accumulator, 81 us for 49 152 rows) stays the fp32 path.
struct BagWgP {
  float* part;     // [splits][N][ldk]
  float* colsum;   // [splits][N] or null
  int R, N, K;     // rows, H, F
  int bn, m_tiles, rows_per_split, stages, b_stages;
  int ldk;         // row pitch of a partial (floats)
  const int64_t* rowptr;
  const int2* ent;
  const float* tail;
  int T, tail_start;
  const int64_t* gather;
};
constexpr int kBagMaxRows = 4096;   // rows of a split whose (row, first entry, count) fit the shared-memory table
constexpr int kBagMaxTiles = 5;     // feature tiles per CTA (F <= 640)

__global__ void __launch_bounds__(kThreadsW, 1)
bag_wgrad_tc_kernel(const __grid_constant__ CUtensorMap tmDy, BagWgP p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int nb_atoms = p.bn >> 5;
  const uint32_t a_bytes = (uint32_t)p.m_tiles * kABytes;        // one A stage: every feature tile of the chunk
  const uint32_t b_bytes = (uint32_t)nb_atoms * kAtomBytes;      // one B stage: the chunk's dh slice
  // two rings: A (p.stages deep, big, filled by the fill warps) and B (p.b_stages deep, small, fed by TMA several chunks ahead)
  uint8_t* ringB = smem + (uint32_t)p.stages * a_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(ringB + (uint32_t)p.b_stages * b_bytes);
  uint64_t* full = bars;                       // [b_stages] dh chunk landed
  uint64_t* ready = bars + kMaxStages;         // [stages]   A expanded and B rounded: the MMAs of the chunk may go
  uint64_t* empty = bars + 2 * kMaxStages;     // [stages]   the MMAs that read the A stage have completed
  uint64_t* acc_full = bars + 3 * kMaxStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3 * kMaxStages + 4);
  uint64_t* emptyB = bars + 4 * kMaxStages;    // [b_stages] the MMAs that read the B stage have completed
  float* tailbuf = reinterpret_cast<float*>(smem);   // epilogue tiles: alias the A ring (free once acc_full has fired)
  int64_t* row_beg = reinterpret_cast<int64_t*>(reinterpret_cast<uint8_t*>(bars) + 512);   // [rows_per_split] first entry of a row
  int32_t* row_g = reinterpret_cast<int32_t*>(row_beg + p.rows_per_split);         // [rows_per_split] feature-matrix row
  int32_t* row_n = row_g + p.rows_per_split;                                       // [rows_per_split] sparse entries of the row

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n0 = blockIdx.x * p.bn;
  const int r_lo = blockIdx.y * p.rows_per_split, r_hi = min(p.R, r_lo + p.rows_per_split);
  const int nchunks = (r_hi - r_lo + kKC - 1) / kKC;
  const bool want_colsum = p.colsum != nullptr;
  const uint32_t need_cols = (uint32_t)(p.m_tiles * p.bn);
  const uint32_t tmem_cols = need_cols <= 32 ? 32u : need_cols <= 64 ? 64u : need_cols <= 128 ? 128u : need_cols <= 256 ? 256u : 512u;
  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmDy);
    for (int i = 0; i < p.stages; ++i) {
      mbar_init(ready + i, kRoundWarps);
      mbar_init(empty + i, 1);
    }
    for (int i = 0; i < p.b_stages; ++i) {
      mbar_init(full + i, 1);
      mbar_init(emptyB + i, 1);
    }
    mbar_init(acc_full, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, tmem_cols);
  // the split's index chain, two global latencies for the whole CTA: idx -> rowptr
  for (int i = threadIdx.x; i < r_hi - r_lo; i += kThreadsW) {
    const int64_t g = p.gather ? p.gather[r_lo + i] : (int64_t)(r_lo + i);
    const int64_t b = p.rowptr[g];
    row_g[i] = (int32_t)g;
    row_beg[i] = b;
    row_n[i] = (int32_t)(p.rowptr[g + 1] - b);
  }
  // the A halves of every stage start as zeros (the fill warps keep them so between chunks)
  for (uint32_t i = threadIdx.x; i < (uint32_t)p.stages * a_bytes / 16; i += kThreadsW)
    reinterpret_cast<uint4*>(smem)[i] = make_uint4(0u, 0u, 0u, 0u);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      for (int c = 0; c < nchunks; ++c) {
        const int bs = c % p.b_stages;
        mbar_wait_relaxed(emptyB + bs, ((c / p.b_stages) & 1) ^ 1, 64);
        uint8_t* sB = ringB + (uint32_t)bs * b_bytes;
        mbar_expect_tx(full + bs, b_bytes);
        const int r = r_lo + c * kKC;
        for (int b = 0; b < nb_atoms; ++b) tma_load_2d(sB + b * kAtomBytes, &tmDy, full + bs, n0 + 32 * b, r);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = make_idesc(2 /*tf32*/, kBM, p.bn, 1, 1);
      for (int c = 0; c < nchunks; ++c) {
        const int stage = c % p.stages, bs = c % p.b_stages;
        mbar_wait(ready + stage, (c / p.stages) & 1);
        tc_fence_after();
        const uint32_t sA = smem_u32(smem + (uint32_t)stage * a_bytes);
        const uint32_t sB = smem_u32(ringB + (uint32_t)bs * b_bytes);
        for (int tile = 0; tile < p.m_tiles; ++tile) {
#pragma unroll
          for (int ks = 0; ks < kKC / 8; ++ks)
            umma_tf32(tmem_base + (uint32_t)(tile * p.bn), make_mnmajor_desc_tf32(sA + tile * kABytes + ks * 1024, kAtomBytes, 512),
                      make_mnmajor_desc_tf32(sB + ks * 1024, kAtomBytes, 512), idesc, (c | ks) != 0 ? 1u : 0u);
        }
        umma_commit(empty + stage);
        umma_commit(emptyB + bs);
      }
      umma_commit(acc_full);
    }
  } else if (warp < 2 + kEpiWarpsW) {
    // ===================== epilogue: transposed store  part[z][n][m]  through a [32][33] tile per warp =====================
    const int quad = warp & 3;
    float* stg = tailbuf + (warp - 2) * (32 * 33);
    float* Cz = p.part + (int64_t)blockIdx.y * p.N * p.ldk;
    mbar_wait(acc_full, 0);
    tc_fence_after();
    for (int tile = 0; tile < p.m_tiles; ++tile) {
      const int m = tile * kBM + quad * 32 + lane;
      for (int col = 0; col < p.bn; col += 32) {
        uint32_t r[32];
        tmem_ld_32x32(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(tile * p.bn + col), r);
#pragma unroll
        for (int i = 0; i < 32; ++i) stg[lane * 33 + i] = __uint_as_float(r[i]);
        __syncwarp();
        const int rows = min(32, p.N - (n0 + col));
        if (m < p.K) {   // lanes = 32 consecutive m: one 128-byte store per output row n
          float* dst = Cz + (int64_t)(n0 + col) * p.ldk + m;
          const float* srow = stg + lane * 33;
#pragma unroll 8
          for (int i = 0; i < 32; ++i)
            if (i < rows) dst[(int64_t)i * p.ldk] = srow[i];
        }
        __syncwarp();
      }
    }
  } else {
    // ===================== fill warps: expand the chunk's CSR rows into the A stage; round dh (+ exact column sums) ==========
    const int t = threadIdx.x - 32 * (2 + kEpiWarpsW);   // 0 .. 127
    const int crow = t >> 2, sub = t & 3;                 // row of the chunk, which quarter of its entries
    float4 csum[8];
#pragma unroll
    for (int b = 0; b < 8; ++b) csum[b] = make_float4(0.f, 0.f, 0.f, 0.f);
    // entries / tail values of THIS thread for one chunk (held one chunk ahead)
    int2 e0 = make_int2(-1, 0), e1 = make_int2(-1, 0);
    float tl0 = 0.f, tl1 = 0.f;
    int n_row = 0;
    int64_t beg_row = 0;
    auto fetch = [&](int c) {
      e0 = e1 = make_int2(-1, 0);
      tl0 = tl1 = 0.f;
      n_row = 0;
      const int i = c * kKC + crow;
      if (c < nchunks && r_lo + i < r_hi) {
        n_row = row_n[i];
        beg_row = row_beg[i];
        if (sub < n_row) e0 = p.ent[beg_row + sub];
        if (sub + 4 < n_row) e1 = p.ent[beg_row + sub + 4];
        const float* ts = p.tail + (int64_t)row_g[i] * p.T;
        if (sub < p.T) tl0 = ts[sub];
        if (sub + 4 < p.T) tl1 = ts[sub + 4];
      }
    };
    // byte offset of element (feature f, chunk row crow) inside the A part of a stage
    auto where = [&](int f) -> uint32_t {
      const int tile = f >> 7, fl = f & 127;
      const int atom = fl >> 5, fi = fl & 31, L = fi >> 2;
      const int ch = ((((L >> 1) ^ (crow & 3)) << 1) | (L & 1));
      return (uint32_t)(tile * (int)kABytes + atom * (int)kAtomBytes + crow * 128 + ch * 16 + (fi & 3) * 4);
    };
    // what this thread scattered into each stage last time (to be zeroed when the stage comes round): 4 words + the slow rows
    uint32_t mine[kMaxStages > 4 ? 4 : kMaxStages][4];
    int slow_n[4];
    int64_t slow_beg[4];
#pragma unroll
    for (int sidx = 0; sidx < 4; ++sidx) {
      slow_n[sidx] = 0; slow_beg[sidx] = 0;
#pragma unroll
      for (int k = 0; k < 4; ++k) mine[sidx][k] = 0xFFFFFFFFu;
    }
    // two chunks ahead: the loads of chunk c+2 go out while chunk c is expanded (one chunk ahead the loop stalled ~0.5 us per
    // chunk on them - globaltimer trace of the first version)
    fetch(0);
    int2 n0e = e0, n1e = e1;
    float nt0 = tl0, nt1 = tl1;
    int nn_row = n_row;
    int64_t nbeg_row = beg_row;
    fetch(1);
    for (int c = 0; c < nchunks; ++c) {
      const int stage = c % p.stages;
      const int2 a0 = n0e, a1 = n1e;
      const float u0 = nt0, u1 = nt1;
      const int n_cur = nn_row;
      const int64_t beg_cur = nbeg_row;
      const bool row_ok = c * kKC + crow < r_hi - r_lo;
      n0e = e0; n1e = e1; nt0 = tl0; nt1 = tl1; nn_row = n_row; nbeg_row = beg_row;   // chunk c+1 (loaded during chunk c-1)
      fetch(c + 2);
      mbar_wait(empty + stage, ((c / p.stages) & 1) ^ 1);   // the MMAs that read this stage last time have completed
      uint8_t* sA = smem + (uint32_t)stage * a_bytes;
      // zeros back into the words this thread set in this stage p.stages chunks ago (nobody else touched them), then this chunk
#pragma unroll
      for (int sidx = 0; sidx < 4; ++sidx) {
        if (sidx == stage) {
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            if (mine[sidx][k] != 0xFFFFFFFFu) *reinterpret_cast<uint32_t*>(sA + mine[sidx][k]) = 0u;
            mine[sidx][k] = 0xFFFFFFFFu;
          }
          for (int k = sub + 8; k < slow_n[sidx]; k += 4) *reinterpret_cast<uint32_t*>(sA + where(p.ent[slow_beg[sidx] + k].x)) = 0u;
          slow_n[sidx] = n_cur; slow_beg[sidx] = beg_cur;
          if (a0.x >= 0 && a0.x < p.K) { mine[sidx][0] = where(a0.x); *reinterpret_cast<uint32_t*>(sA + mine[sidx][0]) = to_tf32(__int_as_float(a0.y)) & 0xFFFFE000u; }
          if (a1.x >= 0 && a1.x < p.K) { mine[sidx][1] = where(a1.x); *reinterpret_cast<uint32_t*>(sA + mine[sidx][1]) = to_tf32(__int_as_float(a1.y)) & 0xFFFFE000u; }
          if (row_ok && sub < p.T) { mine[sidx][2] = where(p.tail_start + sub); *reinterpret_cast<uint32_t*>(sA + mine[sidx][2]) = to_tf32(u0) & 0xFFFFE000u; }
          if (row_ok && sub + 4 < p.T) { mine[sidx][3] = where(p.tail_start + sub + 4); *reinterpret_cast<uint32_t*>(sA + mine[sidx][3]) = to_tf32(u1) & 0xFFFFE000u; }
        }
      }
      for (int k = sub + 8; k < n_cur; k += 4) {            // rows with more than 8 sparse entries (not on the fast path)
        const int2 e = p.ent[beg_cur + k];
        if (e.x >= 0 && e.x < p.K) *reinterpret_cast<uint32_t*>(sA + where(e.x)) = to_tf32(__int_as_float(e.y)) & 0xFFFFE000u;
      }
      // B: round dh to TF32 in place (+ exact column sums for the bias gradient), as in gemm_tf32_tma_wgrad_kernel
      const int bs = c % p.b_stages;
      mbar_wait(full + bs, (c / p.b_stages) & 1);
#pragma unroll
      for (int i0 = 0; i0 < 16; i0 += 4) {      // four 16-byte pieces per round: the loads of a round go out together
        if (i0 < 2 * nb_atoms) {
          uint4* q = reinterpret_cast<uint4*>(ringB + (uint32_t)bs * b_bytes) + t + 128 * i0;
          float4 v[4];
#pragma unroll
          for (int j = 0; j < 4; ++j)
            v[j] = (i0 + j < 2 * nb_atoms) ? *reinterpret_cast<float4*>(q + 128 * j) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            if (i0 + j < 2 * nb_atoms) {
              if (want_colsum) { csum[(i0 + j) >> 1].x += v[j].x; csum[(i0 + j) >> 1].y += v[j].y; csum[(i0 + j) >> 1].z += v[j].z; csum[(i0 + j) >> 1].w += v[j].w; }
              q[128 * j] = make_uint4(to_tf32(v[j].x), to_tf32(v[j].y), to_tf32(v[j].z), to_tf32(v[j].w));
            }
          }
        }
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(ready + stage);
    }
    if (want_colsum) {
      mbar_wait(acc_full, 0);   // every MMA has read the stages: the ring memory is free
      float4* red = reinterpret_cast<float4*>(ringB);   // [128 threads][8 blocks]: the B ring is free (the epilogue uses the A ring)
#pragma unroll
      for (int b = 0; b < 8; ++b) red[t * 8 + b] = csum[b];
      asm volatile("bar.sync 1, %0;" ::"n"(32 * kRoundWarps) : "memory");
      for (int n = t; n < p.bn; n += 32 * kRoundWarps) {
        const int b = n >> 5, want = (n & 31) >> 2, comp = n & 3;
        float sacc = 0.f;
        for (int u = 0; u < 32 * kRoundWarps; ++u) {
          const int cu = ((((u & 7) >> 1) ^ ((u >> 3) & 3)) << 1) | (u & 1);
          if (cu == want) sacc += reinterpret_cast<const float*>(red + u * 8 + b)[comp];
        }
        if (n0 + n < p.N) p.colsum[(int64_t)blockIdx.y * p.N + n0 + n] = sacc;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, tmem_cols);
}

static bool aligned16(const void* p) { return ((uintptr_t)p & 15) == 0; }

}  // namespace tma

using namespace tma;

// Returns TTAM_OK, an error, or +1 when the operands do not qualify for TMA (the caller falls back to gemm_tc.cu).
int tma_gemm(const float* A, int64_t lda, const float* B, int64_t ldb, float* C, int64_t ldc, int64_t M, int64_t N, int64_t K,
             const float* bias, int relu, float dropout_p, uint64_t seed, uint64_t offset, const ttam_step_state* st,
             const float* aux, int64_t ldaux, int mask_mode, float scale, int accumulate, int roundA, int round_out,
             cudaStream_t stream) {
  if (!aligned16(A) || !aligned16(B) || lda % 4 || ldb % 4 || K % 4 || M <= 0 || N <= 0 || K <= 0) return 1;
  if (getenv("TTAM_NO_TMA_GEMM")) return 1;
  TmaP p{};
  p.C = C; p.ldc = ldc; p.M = (int)M; p.N = (int)N; p.K = (int)K;
  // N tile: a multiple of 16 (UMMA N for M = 128), at most 256; wide N is split into equal tiles
  const int n_tiles = (int)ceil_div(N, 256);
  p.bn = (int)align_up(ceil_div(N, n_tiles), 16);
  p.n_tiles_n = (int)ceil_div(N, p.bn);
  p.acc_stride = (int)align_up(p.bn, 32);
  p.units = (int)ceil_div(M, kBM) * p.n_tiles_n;
  p.nchunks = (int)ceil_div(K, kKC);
  const size_t stage_bytes = kABytes + (size_t)p.bn * 128;
  // ring depth: ~128 kB of operands in flight per SM is several times what the HBM latency needs at this CTA's share of the
  // bandwidth, and it leaves room for a CTA of another kernel (the weight-gradient GEMMs run next to the data-gradient chain)
  p.stages = (int)((128 * 1024) / stage_bytes);
  if (const char* e = getenv("TTAM_TMA_STAGES")) p.stages = atoi(e);
  if (p.stages < 3) p.stages = 3;
  if (p.stages > kMaxStages) p.stages = kMaxStages;
  if ((size_t)p.stages * stage_bytes > 196 * 1024) return 1;
  p.roundA = roundA; p.round_out = round_out;
  p.vecC = aligned16(C) && ldc % 4 == 0;
  p.bias = bias; p.vecBias = bias != nullptr && aligned16(bias);
  p.relu = relu; p.dropout_p = dropout_p; p.seed = seed; p.offset = offset; p.st = st;
  p.aux = aux; p.ldaux = ldaux; p.mask_mode = mask_mode; p.vecAux = aux != nullptr && aligned16(aux) && ldaux % 4 == 0;
  p.scale = scale; p.accumulate = accumulate;
  CUtensorMap tmA, tmB;
  int rc = make_tmap_2d(&tmA, A, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, (uint64_t)M, (uint64_t)K, (uint64_t)lda * 4, kKC, kBM,
                        CU_TENSOR_MAP_SWIZZLE_128B);
  if (rc != TTAM_OK) return rc;
  rc = make_tmap_2d(&tmB, B, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, (uint64_t)N, (uint64_t)K, (uint64_t)ldb * 4, kKC, (uint32_t)p.bn,
                    CU_TENSOR_MAP_SWIZZLE_128B);
  if (rc != TTAM_OK) return rc;
  const size_t smem = (size_t)p.stages * stage_bytes + 1024 /*alignment*/ + 512 /*barriers*/ + (size_t)kEpiWarps * 32 * 36 * 4 /*epilogue tiles*/;
  static bool attr_done = false;
  if (!attr_done) {
    TTAM_CUDA(cudaFuncSetAttribute(gemm_tf32_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
    attr_done = true;
  }
  const int grid = p.units < num_sms() ? p.units : num_sms();
  gemm_tf32_tma_kernel<<<grid, kThreads, smem, stream>>>(tmA, tmB, p);
  TTAM_LAUNCH_CHECK();
  return TTAM_OK;
}

// Splits of the TMA weight gradient: about two CTAs per SM, at least 256 rows each.
int tma_wgrad_splits(int64_t M, int64_t N, int64_t K) {
  const int64_t tiles = ceil_div(K, kBM) * ceil_div(N, 256);
  static const int tenths = [] { const char* e = getenv("TTAM_WGRAD_CTAS_X10"); const int v = e ? atoi(e) : 0; return v > 0 ? v : 20; }();
  int64_t s = (tenths * (int64_t)num_sms() / 10) / tiles;   // CTAs per SM x 10 (default 2.0)
  const int64_t by_rows = ceil_div(M, 256);
  if (s > by_rows) s = by_rows;
  return (int)(s < 1 ? 1 : s);
}

// part[z][N][K] (and colsum[z][N]) for z < *real_splits.  +1: operands do not qualify (the caller uses gemm_tc.cu).
int tma_wgrad_partials(const float* dy, int64_t lddy, const float* x, int64_t ldx, float* partial, float* colsum_partial, int64_t M,
                       int64_t N, int64_t K, int* real_splits, int prerounded, cudaStream_t stream) {
  if (!aligned16(dy) || !aligned16(x) || lddy % 4 || ldx % 4 || M <= 0 || N <= 0 || K <= 0) return 1;
  if (getenv("TTAM_NO_TMA_GEMM") || getenv("TTAM_NO_TMA_WGRAD")) return 1;
  WgP p{};
  p.part = partial; p.colsum = colsum_partial; p.R = (int)M; p.N = (int)N; p.K = (int)K;
  const int n_tiles = (int)ceil_div(N, 256);
  p.bn = (int)align_up(ceil_div(N, n_tiles), 32);
  const int nt = (int)ceil_div(N, p.bn);
  p.m_tiles = (int)ceil_div(K, kBM);
  const int splits = tma_wgrad_splits(M, N, K);
  p.rows_per_split = (int)align_up(ceil_div(M, splits), kKC);
  *real_splits = (int)ceil_div(M, p.rows_per_split);
  p.roundA = !(prerounded & 1); p.roundB = 1;
  const size_t stage_bytes = kABytes + (size_t)(p.bn / 32) * kAtomBytes;
  p.stages = (int)((96 * 1024) / stage_bytes);       // two CTAs per SM
  if (p.stages < 2) p.stages = 2;
  if (p.stages > kMaxStages) p.stages = kMaxStages;
  CUtensorMap tmX, tmDy;
  int rc = make_tmap_2d(&tmX, x, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, (uint64_t)M, (uint64_t)K, (uint64_t)ldx * 4, 32, 32,
                        CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B);
  if (rc != TTAM_OK) return rc;
  rc = make_tmap_2d(&tmDy, dy, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, (uint64_t)M, (uint64_t)N, (uint64_t)lddy * 4, 32, 32,
                    CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B);
  if (rc != TTAM_OK) return rc;
  size_t tail = (size_t)kEpiWarpsW * 32 * 33 * 4;
  size_t smem = (size_t)p.stages * stage_bytes;
  if (smem < (size_t)128 * 8 * 16) smem = (size_t)128 * 8 * 16;     // the column-sum reduction reuses the ring
  smem += 1024 + 512 + tail;
  static bool attr_done = false;
  if (!attr_done) {
    TTAM_CUDA(cudaFuncSetAttribute(gemm_tf32_tma_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
    attr_done = true;
  }
  dim3 grid((unsigned)(p.m_tiles * nt), (unsigned)*real_splits);
  gemm_tf32_tma_wgrad_kernel<<<grid, kThreadsW, smem, stream>>>(tmX, tmDy, p);
  TTAM_LAUNCH_CHECK();
  return TTAM_OK;
}

// Cut of the bag-form weight gradient: every CTA owns all feature tiles (m_tiles accumulators of bn TMEM columns), one slice of bn
// hidden columns and one split of the rows; one CTA per SM.
static bool bag_wgrad_cut(int64_t R, int64_t H, int64_t F, int* bn, int* n_tiles, int* splits) {
  const int64_t m_tiles = ceil_div(F, kBM);
  if (m_tiles > kBagMaxTiles || H % 32 || H <= 0) return false;
  int64_t w = (512 / m_tiles) / 32 * 32;
  if (w > 256) w = 256;
  if (w > H) w = H;
  if (w < 32) return false;
  *bn = (int)w;
  *n_tiles = (int)ceil_div(H, w);
  int64_t s = num_sms() / *n_tiles;
  const int64_t by_rows = ceil_div(R, 64);
  if (s > by_rows) s = by_rows;
  if (s < 1) s = 1;
  while (align_up(ceil_div(R, s), kKC) > kBagMaxRows) ++s;
  *splits = (int)s;
  return true;
}
int tma_bag_wgrad_splits(int64_t R, int64_t H, int64_t F) {
  int bn, nt, s;
  return bag_wgrad_cut(R, H, F, &bn, &nt, &s) ? s : 1;
}

// part[z][H][F] (and colsum[z][H]) for z < *real_splits.  +1: shape not covered (the caller keeps the SIMT bag kernel).
int tma_bag_wgrad_partials(const int64_t* rowptr, const void* entries, const float* tail, int64_t T, int64_t tail_start,
                           const int64_t* gather, int64_t R, const float* dh, int64_t lddh, float* partial, float* colsum_partial,
                           int64_t H, int64_t F, int* real_splits, cudaStream_t stream) {
  if (!aligned16(dh) || lddh % 4 || R <= 0 || F <= 0 || T < 0 || T > 8) return 1;
  if (getenv("TTAM_NO_TMA_GEMM") || getenv("TTAM_NO_BAG_TC")) return 1;
  BagWgP p{};
  int n_tiles = 1, splits = 1;
  if (!bag_wgrad_cut(R, H, F, &p.bn, &n_tiles, &splits)) return 1;
  p.part = partial; p.colsum = colsum_partial; p.R = (int)R; p.N = (int)H; p.K = (int)F;
  p.m_tiles = (int)ceil_div(F, kBM);
  p.ldk = (int)F;   // (a 32-float row pitch - aligned 128-byte stores - was measured: no faster, and it needs its own reduce kernel)
  p.rows_per_split = (int)align_up(ceil_div(R, splits), kKC);
  *real_splits = (int)ceil_div(R, p.rows_per_split);
  p.rowptr = rowptr; p.ent = (const int2*)entries; p.tail = tail; p.T = (int)T; p.tail_start = (int)tail_start; p.gather = gather;
  const size_t a_bytes = (size_t)p.m_tiles * kABytes, b_bytes = (size_t)(p.bn / 32) * kAtomBytes;
  const size_t fixed = 1024 + 512 + (size_t)p.rows_per_split * 16;
  p.stages = 2;
  if (p.stages * a_bytes < (size_t)kEpiWarpsW * 32 * 33 * 4) return 1;   // the epilogue tiles alias the A ring
  // the B ring takes what is left (dh chunks several chunks ahead of the MMAs; its 16 kB floor also holds the column-sum reduction)
  int64_t bst = ((int64_t)220 * 1024 - (int64_t)fixed - (int64_t)(p.stages * a_bytes)) / (int64_t)b_bytes;
  if (bst > kMaxStages) bst = kMaxStages;
  if (bst < 2 || (size_t)bst * b_bytes < (size_t)128 * 8 * 16) return 1;
  p.b_stages = (int)bst;
  const size_t smem = (size_t)p.stages * a_bytes + (size_t)p.b_stages * b_bytes + fixed;
  CUtensorMap tmDy;
  int rc = make_tmap_2d(&tmDy, dh, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, (uint64_t)R, (uint64_t)H, (uint64_t)lddh * 4, 32, 32,
                        CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B);
  if (rc != TTAM_OK) return rc;
  static bool attr_done = false;
  if (!attr_done) {
    TTAM_CUDA(cudaFuncSetAttribute(bag_wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
    attr_done = true;
  }
  dim3 grid((unsigned)n_tiles, (unsigned)*real_splits);
  bag_wgrad_tc_kernel<<<grid, kThreadsW, smem, stream>>>(tmDy, p);
  TTAM_LAUNCH_CHECK();
  return TTAM_OK;
}

// ---- per-step weight preparation: TF32-rounded copy and TF32-rounded transposed copy of up to 4 small matrices ------------
struct PrepItem {
  const float* src;
  float* dst;
  float* dst_t;
  int rows, cols;
  int64_t ld;
};
struct PrepList {
  PrepItem it[4];
  int count;
};
__global__ void prepare_weights_kernel(PrepList L) {
  __shared__ float tile[32][33];
  const PrepItem w = L.it[blockIdx.z];
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  if (r0 >= w.rows || c0 >= w.cols) return;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int r = r0 + i, c = c0 + threadIdx.x;
    float v = 0.f;
    if (r < w.rows && c < w.cols) {
      v = tma::round_out_tf32(w.src[(int64_t)r * w.ld + c]);
      if (w.dst) w.dst[(int64_t)r * w.cols + c] = v;
    }
    tile[i][threadIdx.x] = v;
  }
  __syncthreads();
  if (w.dst_t)
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
      const int c = c0 + i, r = r0 + threadIdx.x;
      if (c < w.cols && r < w.rows) w.dst_t[(int64_t)c * w.rows + r] = tile[threadIdx.x][i];
    }
}

}  // namespace ttam

using namespace ttam;

extern "C" int ttam_prepare_weights(const float* const* src, const int64_t* ld, const int64_t* rows, const int64_t* cols,
                                    float* const* dst, float* const* dst_t, int64_t count, void* stream) {
  TTAM_CHECK_ARG(src && ld && rows && cols && dst && dst_t && count >= 1 && count <= 4, "prepare_weights: 1..4 matrices");
  PrepList L{};
  L.count = (int)count;
  int64_t mr = 0, mc = 0;
  for (int i = 0; i < count; ++i) {
    TTAM_CHECK_ARG(src[i] && rows[i] > 0 && cols[i] > 0 && ld[i] >= cols[i], "prepare_weights: bad matrix %d", i);
    L.it[i] = PrepItem{src[i], dst[i], dst_t[i], (int)rows[i], (int)cols[i], ld[i]};
    mr = rows[i] > mr ? rows[i] : mr;
    mc = cols[i] > mc ? cols[i] : mc;
  }
  prepare_weights_kernel<<<dim3((unsigned)ceil_div(mc, 32), (unsigned)ceil_div(mr, 32), (unsigned)count), dim3(32, 8), 0, (cudaStream_t)stream>>>(L);
  TTAM_LAUNCH_CHECK();
  return TTAM_OK;
}
