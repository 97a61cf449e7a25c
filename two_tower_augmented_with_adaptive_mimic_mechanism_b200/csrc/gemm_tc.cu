// TF32 tensor-core GEMMs (tcgen05.mma kind::tf32, fp32 accumulate in TMEM) for the tower linear layers:
// TTAM_PREC_TF32 of ttam_linear_fwd / _dgrad / _wgrad (reference encoders.py:121-144,157-162 + autograd).
//
// One kernel, three roles.  C[M,N] = sum_k A(m,k) B(n,k) with each operand either K-major (the reduction index is
// contiguous in global memory) or MN-major (the row/column index is contiguous):
//   fwd   : A = x[gather(m), k]  K-major        B = w[n, k]          K-major      bias + activation + dropout epilogue
//   dgrad : A = dy[m, n']        K-major        B = w[n', k'] as (k', n')  MN-major   mask + scale (+ accumulate) epilogue
//   wgrad : A = x[gather(r), k'] as (k', r)  MN-major, B = dy[r, n'] as (n', r)  MN-major, reduction over the rows r,
//           split over blockIdx.z, transposed partial store  part[z][n'][k']  (summed by splitk_reduce_kernel)
// Operands stay fp32 in global memory; 8 producer warps copy them with cp.async (16-byte LDGSTS, stages-1 chunks in
// flight per thread, zero-filled tails), round each landed piece to TF32 (nearest) in place, and publish the stage
// into shared memory in the canonical swizzled UMMA layouts (K-major: SWIZZLE_128B; MN-major tf32:
// SWIZZLE_128B_BASE32B, the only one the hardware accepts), 32 reduction steps per stage; one thread
// issues 4 tcgen05.mma (M=128, N=bn, K=8) per stage; the producers then become the epilogue (tcgen05.ld, thread = row).
// The layer-1 GEMMs are bound by L2->SM operand traffic (2.4 kB of features + the 465 kB weight per 128 rows), not by
// the tensor pipe, which is why the operands are not down-converted further.
#include "gemm_tc.cuh"
#include "sm100.cuh"
#include <cstdlib>

namespace ttam {
namespace tcg {

using namespace ttam::sm100;

constexpr int kBM = 128;
constexpr int kKC = 32;  // reduction elements per stage: one 128-byte swizzled row of fp32
constexpr int kProdWarps = 8;
constexpr int kProdThreads = 32 * kProdWarps;
constexpr int kThreads = kProdThreads + 32;
constexpr uint32_t kABytes = kBM * 128;  // 16 KB per stage
constexpr uint32_t kMnLbo = kKC * 128;   // MN-major: byte distance between 32-element atoms along M/N (32 rows x 128 B)

struct TcP {
  const float* A;
  const float* B;
  float* C;
  int64_t lda, ldb, ldc;
  const int64_t* gatherA;  // K-major A: row map (m -> source row); MN-major A: reduction-row map (k -> source row)
  const int64_t* gatherB;  // MN-major B: reduction-row map
  int M, N, K;
  int k_chunk;  // reduction range per blockIdx.z (multiple of kKC); 0 = whole K
  int bn;       // N tile: multiple of 32, <= 256
  int stages;
  int vecA, vecB, vecC, vecAux, vecBias;  // 16-byte accesses allowed (base and leading dimension aligned)
  int roundA, roundB;    // 0: the operand already holds TF32-representable values (rounded when it was laid out)
  int transposed;        // store C(m,n) at C[z][n*ldc + m]
  float* colsum_b;       // MN-major B only: colsum_b[z][n] = sum over this CTA's reduction range of B(n, k)  (bias gradient)
  const float* bias;
  int act;
  float dropout_p;
  uint64_t seed, offset;
  const ttam_step_state* st;
  const float* aux;
  int64_t ldaux;
  int mask_mode;
  float scale;
  int accumulate;
};

// Round-to-nearest (ties away from zero) onto the TF32 grid, as cvt.rna.tf32.f32 does, with integer arithmetic: add
// half a TF32 ulp to the magnitude; the tensor core then drops the low 13 mantissa bits.  (cvt.rna.tf32 compiles to a
// long NaN/Inf-checking sequence: the rounding pass used to cost as much as everything else in the kernel.)  Inf/NaN
// pass through unchanged.
__device__ __forceinline__ uint32_t to_tf32(float x) {
  const uint32_t u = __float_as_uint(x);
  return ((u & 0x7F800000u) != 0x7F800000u) ? u + 0x1000u : u;
}

// ---- asynchronous global -> shared copies (LDGSTS): no registers, many chunks in flight per thread -------------------
// 16-byte copy; bytes past `valid_bytes` (0..16) are zero-filled; valid_bytes == 0 reads nothing
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, int valid_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(valid_bytes) : "memory");
}
__device__ __forceinline__ void cp_async4(uint32_t dst, const void* src, int valid_bytes) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(dst), "l"(src), "r"(valid_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_pending(int n) {  // wait until at most n groups are pending
  switch (n) {
    case 0: asm volatile("cp.async.wait_group 0;" ::: "memory"); break;
    case 1: asm volatile("cp.async.wait_group 1;" ::: "memory"); break;
    case 2: asm volatile("cp.async.wait_group 2;" ::: "memory"); break;
    default: asm volatile("cp.async.wait_group 3;" ::: "memory"); break;
  }
}
// one 16-byte piece: 4 floats starting at src, the first n_valid of them real (aligned source -> one 16-byte copy)
template <bool VEC>
__device__ __forceinline__ void copy_piece(uint32_t dst, const float* src, int n_valid, const float* safe) {
  n_valid = n_valid < 0 ? 0 : (n_valid > 4 ? 4 : n_valid);
  if (n_valid == 0) src = safe;
  if (VEC) {
    cp_async16(dst, src, n_valid * 4);
  } else {
#pragma unroll
    for (int e = 0; e < 4; ++e) cp_async4(dst + 4 * e, e < n_valid ? src + e : safe, e < n_valid ? 4 : 0);
  }
}
// round the 4 floats of a landed piece to TF32 (nearest) in place: the tensor core would truncate otherwise
__device__ __forceinline__ void round_piece(uint8_t* ptr) {
  float4 v = *reinterpret_cast<float4*>(ptr);
  *reinterpret_cast<uint4*>(ptr) = make_uint4(to_tf32(v.x), to_tf32(v.y), to_tf32(v.z), to_tf32(v.w));
}

// VEC: both operands may be copied with 16-byte LDGSTS (bases and leading dimensions 16-byte aligned); the scalar
// instantiation exists for arbitrary layouts and is not the fast path.
template <bool A_MN, bool B_MN, bool VEC>
__global__ void __launch_bounds__(kThreads, 3) gemm_tf32_kernel(TcP p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int bn = p.bn;
  const uint32_t stage_bytes = kABytes + (uint32_t)bn * 128u;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (uint32_t)p.stages * stage_bytes);
  uint64_t* full = bars;
  uint64_t* empty = bars + p.stages;
  uint64_t* acc_full = bars + 2 * p.stages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * p.stages + 1);

  const int t = threadIdx.x;
  const int warp = t >> 5, lane = t & 31;
  const int m0 = blockIdx.y * kBM, n0 = blockIdx.x * bn;
  int k_lo = 0, k_hi = p.K;
  if (p.k_chunk > 0) {
    k_lo = blockIdx.z * p.k_chunk;
    k_hi = min(p.K, k_lo + p.k_chunk);
  }
  const int nchunks = (k_hi - k_lo + kKC - 1) / kKC;
  const uint32_t tmem_cols = bn <= 32 ? 32u : bn <= 64 ? 64u : bn <= 128 ? 128u : 256u;

  if (t == 0) {
#pragma unroll 1
    for (int i = 0; i < p.stages; ++i) {
      mbar_init(full + i, kProdWarps);
      mbar_init(empty + i, 1);
    }
    mbar_init(acc_full, 1);
    fence_barrier_init();
  }
  if (warp == kProdWarps) tmem_alloc(tmem_slot, tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp < kProdWarps) {
    // ===================== producers: cp.async pipeline, `depth` chunks in flight per thread =====================
    // The 16-byte pieces of a stage that THIS thread copies are the same in every chunk: their shared-memory offsets
    // and source pointers are set up once, a chunk costs one LDGSTS + one pointer bump per piece.
    //   K-major tile : piece j = row (t >> 3) + 32 j, 16-byte chunk t & 7 of the row's 128 bytes; source advances 32 floats
    //   MN-major tile: warp w owns reduction rows w + 8 j, lanes walk the 16-byte chunks along M/N (B: two passes h);
    //                  the source row changes with every chunk (k-th row, or gather[k])
    constexpr int kMaxPieces = 12;  // 4 of A + up to 8 of B
    uint32_t doff[kMaxPieces];
    const float* sptr[kMaxPieces];  // K-major: current source (nullptr: row out of range -> zero fill)
    int mn_valid[kMaxPieces];        // MN-major: valid floats of the piece (column range), 0 = piece unused
    const int kc = t & 7, krow = t >> 3;
    const int nb_k = bn >> 5, nb_mn = (bn + 127) >> 7;
#pragma unroll
    for (int i = 0; i < kMaxPieces; ++i) { doff[i] = 0; sptr[i] = nullptr; mn_valid[i] = 0; }
    if (!A_MN) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int row = krow + 32 * j, r = m0 + row;
        doff[j] = (uint32_t)(row * 128 + ((kc ^ (row & 7)) << 4));
        sptr[j] = r < p.M ? p.A + (p.gatherA ? p.gatherA[r] : (int64_t)r) * p.lda + k_lo + 4 * kc : nullptr;
      }
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int m = m0 + 4 * lane;
        doff[j] = mnmajor_tf32_offset(lane >> 3, warp + 8 * j, lane & 7, kMnLbo);
        mn_valid[j] = max(0, min(4, p.M - m));
      }
    }
    if (!B_MN) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int row = krow + 32 * j, n = n0 + row;
        doff[4 + j] = kABytes + (uint32_t)(row * 128 + ((kc ^ (row & 7)) << 4));
        sptr[4 + j] = (j < nb_k && n < p.N) ? p.B + (int64_t)n * p.ldb + k_lo + 4 * kc : nullptr;
      }
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int cc = lane + 32 * h, n = n0 + 4 * cc;
          doff[4 + 2 * j + h] = kABytes + mnmajor_tf32_offset(cc >> 3, warp + 8 * j, cc & 7, kMnLbo);
          mn_valid[4 + 2 * j + h] = (h < nb_mn && 4 * cc < bn) ? max(0, min(4, p.N - n)) : -1;  // -1: no such piece
        }
    }
    int64_t arow[4] = {0, 0, 0, 0};  // MN-major gathered A: source rows of the NEXT chunk to issue
    if (A_MN && p.gatherA) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int kk = k_lo + warp + 8 * j;
        if (kk < k_hi) arow[j] = p.gatherA[kk];
      }
    }
    const int nB = B_MN ? 8 : nb_k;  // B pieces in use (K-major: the first nb_k)
    const bool want_colsum = B_MN && p.colsum_b != nullptr && blockIdx.y == 0;
    float4 csum[2] = {make_float4(0.f, 0.f, 0.f, 0.f), make_float4(0.f, 0.f, 0.f, 0.f)};
    const int depth = p.stages - 1;  // chunk c + depth reuses the stage of chunk c - 1, whose MMAs were issued long ago
    auto issue = [&](int c) {
      const int stage = c % p.stages;
      mbar_wait(empty + stage, ((c / p.stages) & 1) ^ 1);
      const uint32_t base = smem_u32(smem + (uint32_t)stage * stage_bytes);
      const int k0 = k_lo + c * kKC;
      const int k_left = k_hi - (k0 + 4 * kc);  // K-major pieces: valid floats from this thread's chunk position on
      if (!A_MN) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          copy_piece<VEC>(base + doff[j], sptr[j], sptr[j] ? k_left : 0, p.A);
          if (sptr[j]) sptr[j] += kKC;
        }
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int kk = k0 + warp + 8 * j;
          const float* src = nullptr;
          // the gathered source row of this chunk was fetched one chunk ago (arow): the index load is off the issue path
          if (kk < k_hi && mn_valid[j] > 0) src = p.A + (p.gatherA ? arow[j] : (int64_t)kk) * p.lda + m0 + 4 * lane;
          copy_piece<VEC>(base + doff[j], src, src ? mn_valid[j] : 0, p.A);
          const int kn = kk + kKC;
          if (p.gatherA && kn < k_hi) arow[j] = p.gatherA[kn];
        }
      }
      if (!B_MN) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          if (j < nb_k) {
            copy_piece<VEC>(base + doff[4 + j], sptr[4 + j], sptr[4 + j] ? k_left : 0, p.B);
            if (sptr[4 + j]) sptr[4 + j] += kKC;
          }
        }
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int kk = k0 + warp + 8 * j;
          const float* row_ptr = kk < k_hi ? p.B + (p.gatherB ? p.gatherB[kk] : (int64_t)kk) * p.ldb + n0 : nullptr;
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const int i = 4 + 2 * j + h;
            if (mn_valid[i] >= 0) {
              const bool ok = row_ptr && mn_valid[i] > 0;
              copy_piece<VEC>(base + doff[i], ok ? row_ptr + 4 * (lane + 32 * h) : nullptr, ok ? mn_valid[i] : 0, p.B);
            }
          }
        }
      }
    };
    for (int c = 0; c < depth; ++c) {
      if (c < nchunks) issue(c);
      cp_async_commit();
    }
    const bool any_round = p.roundA || p.roundB || want_colsum;
    for (int c = 0; c < nchunks; ++c) {
      if (c + depth < nchunks) issue(c + depth);
      cp_async_commit();               // (possibly empty) group: keeps one group per chunk
      cp_async_wait_pending(depth);    // chunk c has landed (this thread's pieces)
      const int stage = c % p.stages;
      if (any_round) {
        // round the landed pieces to TF32 (nearest) in place - the tensor core would truncate - unless the operand was
        // rounded when it was laid out (feature matrix, padded weight copy)
        uint8_t* base = smem + (uint32_t)stage * stage_bytes;
        if (p.roundA) {
#pragma unroll
          for (int j = 0; j < 4; ++j) round_piece(base + doff[j]);
        }
        if (p.roundB || want_colsum) {
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            if (j < nB && (!B_MN || mn_valid[4 + j] >= 0)) {
              float4 v = *reinterpret_cast<float4*>(base + doff[4 + j]);
              if (B_MN && want_colsum) {  // exact fp32 column sums of B (before the rounding): piece 2 j' + h -> pass h
                csum[j & 1].x += v.x; csum[j & 1].y += v.y; csum[j & 1].z += v.z; csum[j & 1].w += v.w;
              }
              if (p.roundB)
                *reinterpret_cast<uint4*>(base + doff[4 + j]) = make_uint4(to_tf32(v.x), to_tf32(v.y), to_tf32(v.z), to_tf32(v.w));
            }
          }
        }
      }
      fence_proxy_async_smem();  // writes of this thread (cp.async data it waited for, rounded pieces) -> async proxy
      __syncwarp();
      if (lane == 0) mbar_arrive(full + stage);
    }

    // ===================== epilogue: thread = output row (TMEM lane), 16 columns per tcgen05.ld =====================
    mbar_wait(acc_full, 0);
    tc_fence_after();
    if (want_colsum) {
      // every MMA has completed: the stage memory is free.  Warp w summed the reduction rows w, w+8, ...; combine the
      // eight partial sums in warp order (deterministic) and write this CTA's share of the bias gradient.
      float* red = reinterpret_cast<float*>(smem);  // [kProdWarps][256]
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int cc = lane + 32 * h;
        if (4 * cc < bn) *reinterpret_cast<float4*>(red + warp * 256 + 4 * cc) = csum[h];
      }
      asm volatile("bar.sync 1, %0;" ::"n"(kProdThreads) : "memory");
      if (t < bn && n0 + t < p.N) {
        float sacc = 0.f;
#pragma unroll
        for (int w = 0; w < kProdWarps; ++w) sacc += red[w * 256 + t];
        p.colsum_b[(int64_t)blockIdx.z * p.N + n0 + t] = sacc;
      }
    }
    const int quad = warp & 3, half = warp >> 2;
    const int m = m0 + quad * 32 + lane;
    const int cols_half = bn >> 1;
    const bool relu = p.act == TTAM_ACT_RELU;
    const bool drop = p.dropout_p > 0.f;
    const float keep_scale = drop ? 1.f / (1.f - p.dropout_p) : 1.f;
    const uint64_t rng_base = p.offset + ((drop && p.st) ? p.st->rng_offset : 0ull);
    float* Cz = p.C;
    if (p.k_chunk > 0) Cz += (int64_t)blockIdx.z * (p.transposed ? (int64_t)p.N * p.ldc : (int64_t)p.M * p.ldc);
    for (int cb = 0; cb < cols_half; cb += 16) {
      const int col = half * cols_half + cb;
      uint32_t r[16];
      tmem_ld_32x16(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)col, r);
      if (m < p.M) {
        float v[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
        if (p.bias) {
          if (p.vecBias && n0 + col + 15 < p.N) {
#pragma unroll
            for (int i = 0; i < 16; i += 4) {
              const float4 b4 = ld_f4(p.bias + n0 + col + i);
              v[i] += b4.x; v[i + 1] += b4.y; v[i + 2] += b4.z; v[i + 3] += b4.w;
            }
          } else {
#pragma unroll
            for (int i = 0; i < 16; ++i)
              if (n0 + col + i < p.N) v[i] += p.bias[n0 + col + i];
          }
        }
        if (relu) {
#pragma unroll
          for (int i = 0; i < 16; ++i) v[i] = fmaxf(v[i], 0.f);
        }
        if (drop) {
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const int n = n0 + col + i;
            if (n < p.N) v[i] = dropout_keep(p.seed, rng_base + (uint64_t)m * (uint64_t)p.N + (uint64_t)n, p.dropout_p) ? v[i] * keep_scale : 0.f;
          }
        }
        if (p.mask_mode == 1) {
          const float* ax = p.aux + (int64_t)m * p.ldaux + n0 + col;
          if (p.vecAux && n0 + col + 15 < p.N) {  // 64 contiguous bytes of this thread's row: four 16-byte loads
#pragma unroll
            for (int i = 0; i < 16; i += 4) {
              const float4 a4 = ld_f4(ax + i);
              v[i] = a4.x > 0.f ? v[i] : 0.f; v[i + 1] = a4.y > 0.f ? v[i + 1] : 0.f;
              v[i + 2] = a4.z > 0.f ? v[i + 2] : 0.f; v[i + 3] = a4.w > 0.f ? v[i + 3] : 0.f;
            }
          } else {
#pragma unroll
            for (int i = 0; i < 16; ++i)
              if (n0 + col + i < p.N) v[i] = (ax[i] > 0.f) ? v[i] : 0.f;
          }
        }
        if (p.scale != 1.f) {
#pragma unroll
          for (int i = 0; i < 16; ++i) v[i] *= p.scale;
        }
        if (p.transposed) {
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const int n = n0 + col + i;
            if (n < p.N) {
              float* dst = Cz + (int64_t)n * p.ldc + m;
              *dst = p.accumulate ? *dst + v[i] : v[i];
            }
          }
        } else {
          float* dst = Cz + (int64_t)m * p.ldc + n0 + col;
          if (p.vecC && n0 + col + 15 < p.N) {
#pragma unroll
            for (int i = 0; i < 16; i += 4) {
              float4 o = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
              if (p.accumulate) {
                const float4 old = ld_f4(dst + i);
                o.x += old.x; o.y += old.y; o.z += old.z; o.w += old.w;
              }
              st_f4(dst + i, o);
            }
          } else {
#pragma unroll
            for (int i = 0; i < 16; ++i)
              if (n0 + col + i < p.N) dst[i] = p.accumulate ? dst[i] + v[i] : v[i];
          }
        }
      }
      __syncwarp();  // tcgen05.ld is warp-collective: reconverge before the next one
    }
  } else if (lane == 0) {
    // ===================== MMA issuer =====================
    const uint32_t idesc = make_idesc(2 /*tf32*/, kBM, bn, A_MN ? 1 : 0, B_MN ? 1 : 0);
    for (int c = 0; c < nchunks; ++c) {
      const int stage = c % p.stages;
      mbar_wait(full + stage, (c / p.stages) & 1);
      tc_fence_after();
      const uint32_t sA = smem_u32(smem + (uint32_t)stage * stage_bytes);
      const uint32_t sB = sA + kABytes;
#pragma unroll
      for (int ks = 0; ks < kKC / 8; ++ks) {
        // one MMA consumes 8 reduction steps: 32 bytes of a K-major row, or two 4-row groups (1024 B) of an MN-major tile
        const uint64_t ad = A_MN ? make_mnmajor_desc_tf32(sA + ks * 1024, kMnLbo, 512) : make_kmajor_desc<128>(sA + ks * 32);
        const uint64_t bd = B_MN ? make_mnmajor_desc_tf32(sB + ks * 1024, kMnLbo, 512) : make_kmajor_desc<128>(sB + ks * 32);
        umma_tf32(tmem_base, ad, bd, idesc, (c | ks) != 0 ? 1u : 0u);
      }
      umma_commit(empty + stage);
    }
    umma_commit(acc_full);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == kProdWarps) tmem_dealloc(tmem_base, tmem_cols);
}

static bool aligned16(const void* p) { return ((uintptr_t)p & 15) == 0; }

static int pick_bn(int64_t N) {
  int64_t bn = N < 256 ? N : 256;
  if (const char* e = getenv("TTAM_TC_BN_SPLIT")) {  // experiment: two narrower N tiles (more CTAs per SM) instead of one wide one
    if (atoi(e) && bn > 128) bn = (bn + 1) / 2;
  }
  return (int)align_up(bn, 32);
}

template <bool A_MN, bool B_MN, bool VEC>
static int launch_v(TcP& p, int splits, cudaStream_t st) {
  // several CTAs per SM (the epilogue of one overlaps the copies of the others; 384 tiles of 49 152 rows fit in one wave
  // at three per SM): bn <= 128 -> 128 TMEM columns and 2 x 28-32 kB of stages each, three CTAs; wider tiles need 256
  // TMEM columns each, two CTAs
  p.stages = 2;
  if (const char* e = getenv("TTAM_TC_STAGES")) {
    const int v = atoi(e);
    if (v >= 2 && v <= 4 && (size_t)v * (kABytes + (size_t)p.bn * 128) + 1280 <= 200 * 1024) p.stages = v;
  }
  const size_t smem = (size_t)p.stages * (kABytes + (size_t)p.bn * 128) + 1024 + 256;
  static bool attr_done = false;
  if (!attr_done) {
    TTAM_CUDA(cudaFuncSetAttribute(gemm_tf32_kernel<A_MN, B_MN, VEC>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    attr_done = true;
  }
  dim3 grid((unsigned)ceil_div(p.N, p.bn), (unsigned)ceil_div(p.M, kBM), (unsigned)splits);
  gemm_tf32_kernel<A_MN, B_MN, VEC><<<grid, kThreads, smem, st>>>(p);
  TTAM_LAUNCH_CHECK();
  return TTAM_OK;
}

template <bool A_MN, bool B_MN>
static int launch(TcP& p, int splits, cudaStream_t st) {
  return (p.vecA && p.vecB) ? launch_v<A_MN, B_MN, true>(p, splits, st) : launch_v<A_MN, B_MN, false>(p, splits, st);
}

}  // namespace tcg

using namespace tcg;

int tc_linear_fwd(const float* x, int64_t ldx, const int64_t* gather, const float* w, int64_t ldw, const float* bias,
                  float* y, int64_t ldy, int64_t M, int64_t N, int64_t K, int act, float dropout_p, uint64_t seed,
                  uint64_t offset, const ttam_step_state* state_dev, int prerounded, cudaStream_t st) {
  if (act != TTAM_ACT_NONE && act != TTAM_ACT_RELU) {
    set_error("linear_fwd: the tensor-core path fuses ReLU only; run other activations with ttam_act_fwd");
    return TTAM_EUNSUPPORTED;
  }
  TcP p{};
  p.A = x; p.B = w; p.C = y; p.lda = ldx; p.ldb = ldw; p.ldc = ldy; p.gatherA = gather;
  p.M = (int)M; p.N = (int)N; p.K = (int)K; p.bn = pick_bn(N);
  p.vecA = aligned16(x) && ldx % 4 == 0; p.vecB = aligned16(w) && ldw % 4 == 0; p.vecC = aligned16(y) && ldy % 4 == 0;
  p.bias = bias; p.act = act; p.dropout_p = dropout_p; p.seed = seed; p.offset = offset; p.st = state_dev; p.scale = 1.f;
  p.vecBias = bias != nullptr && aligned16(bias);
  p.roundA = !(prerounded & 1); p.roundB = !(prerounded & 2);
  return launch<false, false>(p, 1, st);
}

int tc_linear_dgrad(const float* dy, int64_t lddy, const float* w, float* dx, int64_t lddx, const float* aux, int64_t ldaux,
                    int mask_mode, float scale, int accumulate, int64_t M, int64_t N, int64_t K, cudaStream_t st) {
  // dx[M,K] = dy[M,N] . w[N,K]:  reduction over N;  B(k', n') = w[n'*K + k'] is MN-major
  TcP p{};
  p.A = dy; p.B = w; p.C = dx; p.lda = lddy; p.ldb = K; p.ldc = lddx;
  p.M = (int)M; p.N = (int)K; p.K = (int)N; p.bn = pick_bn(K);
  p.vecA = aligned16(dy) && lddy % 4 == 0; p.vecB = aligned16(w) && K % 4 == 0; p.vecC = aligned16(dx) && lddx % 4 == 0;
  p.aux = aux; p.ldaux = ldaux; p.mask_mode = mask_mode; p.scale = scale; p.accumulate = accumulate;
  p.vecAux = aux != nullptr && aligned16(aux) && ldaux % 4 == 0;
  p.roundA = 1; p.roundB = 1;
  return launch<false, true>(p, 1, st);
}

int tc_wgrad_splits(int64_t M, int64_t N, int64_t K) {
  const int64_t tiles = ceil_div(K, kBM) * ceil_div(N, pick_bn(N));
  const int ctas_per_sm = pick_bn(N) <= 128 ? 3 : 2;  // what fits (TMEM columns, shared memory): fill exactly one wave
  int64_t s = ((int64_t)num_sms() * ctas_per_sm) / tiles;
  const int64_t by_rows = ceil_div(M, 256);
  if (s > by_rows) s = by_rows;
  if (s < 1) s = 1;
  return (int)s;
}

int tc_linear_wgrad_partials(const float* dy, int64_t lddy, const float* x, int64_t ldx, const int64_t* gather, float* partial,
                             float* colsum_partial, int64_t M, int64_t N, int64_t K, int* real_splits, int prerounded,
                             cudaStream_t st) {
  // part[z][n'][k'] = sum_{r in chunk z} dy[r, n'] x[g(r), k']:  C(m = k', n = n'), both operands MN-major over r
  const int splits = tc_wgrad_splits(M, N, K);
  const int chunk = (int)align_up(ceil_div(M, splits), kKC);
  *real_splits = (int)ceil_div(M, chunk);
  TcP p{};
  p.A = x; p.B = dy; p.C = partial; p.lda = ldx; p.ldb = lddy; p.ldc = K; p.gatherA = gather;
  p.M = (int)K; p.N = (int)N; p.K = (int)M; p.k_chunk = chunk; p.bn = pick_bn(N);
  p.vecA = aligned16(x) && ldx % 4 == 0; p.vecB = aligned16(dy) && lddy % 4 == 0; p.vecC = 0;
  p.transposed = 1; p.scale = 1.f; p.colsum_b = colsum_partial;
  p.roundA = !(prerounded & 1); p.roundB = 1;   // A = x (bit 0 of prerounded), B = dy
  return launch<true, true>(p, *real_splits, st);
}

}  // namespace ttam
