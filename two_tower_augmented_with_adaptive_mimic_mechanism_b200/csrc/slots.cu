// Fixed-capacity slot route of the row-sharded step (all-to-all #1..#3 with static shapes).
//   ttam_slot_plan   : bucket a rank's requested row ids by owner (id % W) into W x cap slots, stable inside a bucket;
//                      padding slots repeat the bucket's first id; flag = a bucket overflowed cap or is empty.
//   ttam_slot_unpack : requester side of the forward exchange.  Reads the [t | q] rows of its slots from one base
//                      pointer PER OWNER (the local receive buffer of an NCCL all-to-all, or the owners' buffers
//                      themselves through NVLink peer mappings) and writes t, q and o = t + q in request order.
//   ttam_slot_pack   : requester side of the backward exchange.  Writes the gradient rows [a | b] of every slot to one
//                      base pointer PER OWNER (local send buffer, or the owners' receive buffers over NVLink); padding
//                      slots get zeros, so the owner's segment sums and weight gradients see x + 0.
// HBM/NVLink-bound row copies: 16-byte accesses, 4 independent loads in flight per thread.
#include "common.cuh"

namespace ttam {

constexpr int SLOT_BLOCK = 256;        // threads
constexpr int SLOT_PER_BLOCK = 1024;   // ids per block (4 sub-chunks of 256 consecutive ids: keeps the order stable)
constexpr int SLOT_MAX_W = 16;

struct slot_ptrs {
  const float* a[SLOT_MAX_W];
  const float* b[SLOT_MAX_W];
};
struct slot_dsts {
  float* a[SLOT_MAX_W];
  float* b[SLOT_MAX_W];
};

// ---- plan ------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(SLOT_BLOCK) slot_hist_kernel(const int64_t* __restrict__ idx, int64_t R, int W,
                                                               int32_t* __restrict__ block_hist) {
  __shared__ int32_t h[SLOT_MAX_W];
  if (threadIdx.x < SLOT_MAX_W) h[threadIdx.x] = 0;
  __syncthreads();
  const int64_t base = (int64_t)blockIdx.x * SLOT_PER_BLOCK;
#pragma unroll
  for (int k = 0; k < SLOT_PER_BLOCK / SLOT_BLOCK; ++k) {
    const int64_t i = base + k * SLOT_BLOCK + threadIdx.x;
    if (i < R) atomicAdd(&h[(int)(idx[i] % W)], 1);
  }
  __syncthreads();
  if (threadIdx.x < W) block_hist[(int64_t)blockIdx.x * W + threadIdx.x] = h[threadIdx.x];
}

// one warp per owner: exclusive scan of that owner's per-block counts; flag
__global__ void __launch_bounds__(32 * SLOT_MAX_W) slot_scan_kernel(int32_t* __restrict__ block_hist, int nblk, int W,
                                                                    int64_t cap, int32_t* __restrict__ counts,
                                                                    int32_t* __restrict__ flag) {
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  __shared__ int bad;
  if (threadIdx.x == 0) bad = 0;
  __syncthreads();
  if (w < W) {
    int run = 0;
    for (int b0 = 0; b0 < nblk; b0 += 32) {
      const int b = b0 + lane;
      const int c = b < nblk ? block_hist[(int64_t)b * W + w] : 0;
      int inc = c;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int n = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += n;
      }
      if (b < nblk) block_hist[(int64_t)b * W + w] = run + inc - c;   // in place: count -> offset
      run += __shfl_sync(0xffffffffu, inc, 31);
    }
    if (lane == 0) {
      counts[w] = run;
      if (run > cap || run == 0) atomicOr(&bad, 1);
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) *flag = bad;
}

__global__ void __launch_bounds__(SLOT_BLOCK) slot_place_kernel(const int64_t* __restrict__ idx, int64_t R, int W,
                                                                int64_t cap, const int32_t* __restrict__ block_off,
                                                                int64_t* __restrict__ send_idx,
                                                                int64_t* __restrict__ slot_of,
                                                                int32_t* __restrict__ req_of,
                                                                int64_t* __restrict__ first) {
  __shared__ int32_t run[SLOT_MAX_W];                       // ids of each owner placed by earlier sub-chunks
  __shared__ int32_t wcnt[SLOT_BLOCK / 32][SLOT_MAX_W + 1];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x < SLOT_MAX_W) run[threadIdx.x] = threadIdx.x < W ? block_off[(int64_t)blockIdx.x * W + threadIdx.x] : 0;
  const int64_t base = (int64_t)blockIdx.x * SLOT_PER_BLOCK;
  const int64_t n_slots = (int64_t)W * cap;
  for (int k = 0; k < SLOT_PER_BLOCK / SLOT_BLOCK; ++k) {
    const int64_t i = base + k * SLOT_BLOCK + threadIdx.x;
    const int64_t id = i < R ? idx[i] : 0;
    const int o = i < R ? (int)(id % W) : W;                // W = "no id": its own match group, never placed
    const unsigned peers = __match_any_sync(0xffffffffu, o);
    const int rank_w = __popc(peers & ((1u << lane) - 1u));
    if (lane < SLOT_MAX_W + 1) wcnt[warp][lane] = 0;
    __syncwarp();
    if (rank_w == 0) wcnt[warp][o] = __popc(peers);
    __syncthreads();                                        // run[] of the previous round and wcnt[] are visible
    if (o < W) {
      int j = run[o] + rank_w;
      for (int pw = 0; pw < warp; ++pw) j += wcnt[pw][o];
      if (j < cap) {
        const int64_t s = (int64_t)o * cap + j;
        send_idx[s] = id;
        req_of[s] = (int32_t)i;
        slot_of[i] = s;
        if (j == 0) first[o] = id;
      } else {
        // does not fit: the step takes the dynamic route (flag).  The forward half of the step may already be running when
        // the host learns that (ShardedEngine: speculative replay); every consumer of slot_of skips this marker.
        slot_of[i] = n_slots;
      }
    }
    __syncthreads();
    if (threadIdx.x < W) {
      int tot = 0;
#pragma unroll
      for (int pw = 0; pw < SLOT_BLOCK / 32; ++pw) tot += wcnt[pw][threadIdx.x];
      run[threadIdx.x] += tot;
    }
    __syncthreads();                                        // wcnt[] is rewritten by the next round
  }
}

__global__ void __launch_bounds__(256) slot_pad_kernel(int W, int64_t cap, const int32_t* __restrict__ counts,
                                                       const int64_t* __restrict__ first,
                                                       int64_t* __restrict__ send_idx, int32_t* __restrict__ req_of) {
  const int64_t n_slots = (int64_t)W * cap;
  for (int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; s < n_slots; s += (int64_t)gridDim.x * blockDim.x) {
    const int o = (int)(s / cap);
    const int64_t j = s - (int64_t)o * cap;
    const int64_t cnt = counts[o];
    if (j >= cnt) {
      // padding repeats REAL ids of the bucket (zero gradient rows: the owner's touched-row sets stay exact), cycling through
      // them: repeating only the first one handed the owner one segment of ~cap - count duplicates, which the long-segment
      // kernels of the row-wise optimisers then summed serially (85-96 us per table at N = 2)
      // (an EMPTY bucket raises the flag too; until the host has seen it the slots must hold a row its owner has: row o)
      send_idx[s] = cnt > 0 ? send_idx[(int64_t)o * cap + (j - cnt) % (cnt < cap ? cnt : cap)] : (int64_t)o;
      req_of[s] = -1;
    }
  }
}

// ---- unpack ----------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) slot_unpack_kernel(slot_ptrs src, int64_t ld, int W, int64_t cap,
                                                          const int64_t* __restrict__ slot_of, int64_t R, int64_t D,
                                                          float* __restrict__ t_out, float* __restrict__ q_out,
                                                          float* __restrict__ o_out) {
  const int64_t cpr = D >> 2, total = R * cpr, stride = (int64_t)gridDim.x * blockDim.x;
  const int64_t n_slots = (int64_t)W * cap;
  const bool has_q = src.b[0] != nullptr;
  for (int64_t i0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i0 < total; i0 += 4 * stride) {
    float4 vt[4], vq[4];
    int64_t off[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int64_t i = i0 + u * stride;
      vt[u] = vq[u] = make_float4(0.f, 0.f, 0.f, 0.f);
      off[u] = -1;
      if (i < total) {
        const int64_t r = i / cpr, c = i - r * cpr;
        off[u] = r * D + c * 4;
        const int64_t s = slot_of[r];
        if (s < n_slots) {
          const int o = (int)(s / cap);
          const int64_t e = (s - (int64_t)o * cap) * ld + c * 4;
          vt[u] = ld_f4_stream(src.a[o] + e);
          if (has_q) vq[u] = ld_f4_stream(src.b[o] + e);
        }
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      if (off[u] < 0) continue;
      if (t_out) st_f4(t_out + off[u], vt[u]);
      if (q_out) st_f4(q_out + off[u], vq[u]);
      if (o_out) st_f4(o_out + off[u], make_float4(vt[u].x + vq[u].x, vt[u].y + vq[u].y, vt[u].z + vq[u].z, vt[u].w + vq[u].w));
    }
  }
}

// ---- pack ------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) slot_pack_kernel(const float* __restrict__ a, const float* __restrict__ b0,
                                                        int64_t n0, const float* __restrict__ b1,
                                                        const int32_t* __restrict__ req_of, int W, int64_t cap,
                                                        int64_t D, slot_dsts dst, int64_t ld) {
  const int64_t cpr = D >> 2, n_slots = (int64_t)W * cap, total = n_slots * cpr, stride = (int64_t)gridDim.x * blockDim.x;
  const bool has_b = dst.b[0] != nullptr;
  for (int64_t i0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i0 < total; i0 += 4 * stride) {
    float4 va[4], vb[4];
    int64_t s_[4], c_[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int64_t i = i0 + u * stride;
      va[u] = vb[u] = make_float4(0.f, 0.f, 0.f, 0.f);
      s_[u] = -1;
      if (i < total) {
        const int64_t s = i / cpr, c = i - s * cpr;
        s_[u] = s; c_[u] = c;
        const int64_t r = req_of[s];
        if (r >= 0) {
          va[u] = ld_f4(a + r * D + c * 4);
          if (has_b) vb[u] = ld_f4((r < n0 ? b0 : b1) + r * D + c * 4);
        }
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      if (s_[u] < 0) continue;
      const int o = (int)(s_[u] / cap);
      const int64_t e = (s_[u] - (int64_t)o * cap) * ld + c_[u] * 4;
      st_f4(dst.a[o] + e, va[u]);
      if (has_b) st_f4(dst.b[o] + e, vb[u]);
    }
  }
}

// ---- id exchange over peer mappings: what an all-to-all of send_idx would deliver, plus the local row (id / W) ----
struct slot_id_srcs {
  const int64_t* p[SLOT_MAX_W];
};
__global__ void __launch_bounds__(256) slot_ids_kernel(slot_id_srcs src, int W, int64_t cap, int64_t* __restrict__ recv_idx,
                                                       int64_t* __restrict__ local_rows) {
  const int64_t n_slots = (int64_t)W * cap;
  for (int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; s < n_slots; s += (int64_t)gridDim.x * blockDim.x) {
    const int w = (int)(s / cap);
    const int64_t id = src.p[w][s - (int64_t)w * cap];
    recv_idx[s] = id;
    local_rows[s] = id / W;
  }
}

static inline int copy_grid(int64_t chunks) {
  int64_t g = ceil_div(chunks, 256 * 4);
  const int64_t cap = (int64_t)num_sms() * 8;
  return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

}  // namespace ttam

using namespace ttam;

extern "C" int64_t ttam_slot_plan_workspace_bytes(int64_t R, int64_t world) {
  const int64_t nblk = ceil_div(R > 0 ? R : 1, SLOT_PER_BLOCK);
  return align_up(nblk * world * 4, 16) + align_up(world * 4, 16) + world * 8;
}

extern "C" int ttam_slot_plan(const int64_t* idx, int64_t R, int64_t world, int64_t cap, int64_t* send_idx,
                              int64_t* slot_of, int32_t* req_of, int32_t* flag, void* workspace,
                              int64_t workspace_bytes, void* stream) {
  TTAM_CHECK_ARG(world >= 1 && world <= SLOT_MAX_W, "slot_plan: world=%lld not in [1,%d]", (long long)world, SLOT_MAX_W);
  TTAM_CHECK_ARG(R >= 1 && cap >= 1 && R < (1ll << 31) && world * cap < (1ll << 31), "slot_plan: bad R=%lld cap=%lld",
                 (long long)R, (long long)cap);
  TTAM_CHECK_ARG(idx && send_idx && slot_of && req_of && flag && workspace, "slot_plan: null pointer");
  TTAM_CHECK_ARG(workspace_bytes >= ttam_slot_plan_workspace_bytes(R, world), "slot_plan: workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  const int nblk = (int)ceil_div(R, SLOT_PER_BLOCK), W = (int)world;
  char* ws = (char*)workspace;
  int32_t* block_hist = (int32_t*)ws;
  int32_t* counts = (int32_t*)(ws + align_up((int64_t)nblk * W * 4, 16));
  int64_t* first = (int64_t*)(ws + align_up((int64_t)nblk * W * 4, 16) + align_up(W * 4, 16));
  slot_hist_kernel<<<nblk, SLOT_BLOCK, 0, st>>>(idx, R, W, block_hist);
  TTAM_LAUNCH_CHECK();
  slot_scan_kernel<<<1, 32 * SLOT_MAX_W, 0, st>>>(block_hist, nblk, W, cap, counts, flag);
  TTAM_LAUNCH_CHECK();
  slot_place_kernel<<<nblk, SLOT_BLOCK, 0, st>>>(idx, R, W, cap, block_hist, send_idx, slot_of, req_of, first);
  TTAM_LAUNCH_CHECK();
  const int64_t n_slots = world * cap;
  slot_pad_kernel<<<(int)std::min<int64_t>(ceil_div(n_slots, 256), 1184), 256, 0, st>>>(W, cap, counts, first, send_idx, req_of);
  TTAM_LAUNCH_CHECK();
  return TTAM_OK;
}

extern "C" int ttam_slot_unpack(const float* const* t_src, const float* const* q_src, int64_t ld_src, int64_t world,
                                int64_t cap, const int64_t* slot_of, int64_t R, int64_t D, float* t_out, float* q_out,
                                float* o_out, void* stream) {
  TTAM_CHECK_ARG(world >= 1 && world <= SLOT_MAX_W, "slot_unpack: world=%lld not in [1,%d]", (long long)world, SLOT_MAX_W);
  TTAM_CHECK_ARG(t_src && slot_of && R >= 1 && D >= 4 && D % 4 == 0 && ld_src % 4 == 0 && ld_src >= D,
                 "slot_unpack: bad arguments (D=%lld ld=%lld; rows must be 16-byte multiples)", (long long)D, (long long)ld_src);
  TTAM_CHECK_ARG(t_out || o_out, "slot_unpack: no output");
  slot_ptrs p;
  for (int w = 0; w < SLOT_MAX_W; ++w) {
    p.a[w] = w < world ? t_src[w] : nullptr;
    p.b[w] = (q_src && w < world) ? q_src[w] : nullptr;
    TTAM_CHECK_ARG(w >= world || (p.a[w] && ((uintptr_t)p.a[w] & 15) == 0 && ((uintptr_t)p.b[w] & 15) == 0),
                   "slot_unpack: source %d null or not 16-byte aligned", w);
  }
  slot_unpack_kernel<<<copy_grid(R * (D / 4)), 256, 0, (cudaStream_t)stream>>>(p, ld_src, (int)world, cap, slot_of, R, D,
                                                                               t_out, q_out, o_out);
  TTAM_LAUNCH_CHECK();
  return TTAM_OK;
}

extern "C" int ttam_slot_pack(const float* a, const float* b0, int64_t n0, const float* b1, const int32_t* req_of,
                              int64_t world, int64_t cap, int64_t D, float* const* a_dst, float* const* b_dst,
                              int64_t ld_dst, void* stream) {
  TTAM_CHECK_ARG(world >= 1 && world <= SLOT_MAX_W, "slot_pack: world=%lld not in [1,%d]", (long long)world, SLOT_MAX_W);
  TTAM_CHECK_ARG(a && req_of && a_dst && cap >= 1 && D >= 4 && D % 4 == 0 && ld_dst % 4 == 0 && ld_dst >= D,
                 "slot_pack: bad arguments (D=%lld ld=%lld)", (long long)D, (long long)ld_dst);
  TTAM_CHECK_ARG(!b_dst || (b0 || n0 == 0), "slot_pack: b rows missing");
  TTAM_CHECK_ARG(!b_dst || b1 || b0, "slot_pack: b rows missing");
  slot_dsts d;
  for (int w = 0; w < SLOT_MAX_W; ++w) {
    d.a[w] = w < world ? a_dst[w] : nullptr;
    d.b[w] = (b_dst && w < world) ? b_dst[w] : nullptr;
    TTAM_CHECK_ARG(w >= world || (d.a[w] && ((uintptr_t)d.a[w] & 15) == 0 && ((uintptr_t)d.b[w] & 15) == 0),
                   "slot_pack: destination %d null or not 16-byte aligned", w);
  }
  if (b_dst && !b1) b1 = b0;
  if (b_dst && !b0) { b0 = b1; }
  slot_pack_kernel<<<copy_grid(world * cap * (D / 4)), 256, 0, (cudaStream_t)stream>>>(a, b0, n0, b1, req_of, (int)world,
                                                                                        cap, D, d, ld_dst);
  TTAM_LAUNCH_CHECK();
  return TTAM_OK;
}

extern "C" int ttam_slot_ids(const int64_t* const* src, int64_t world, int64_t cap, int64_t* recv_idx, int64_t* local_rows,
                             void* stream) {
  TTAM_CHECK_ARG(world >= 1 && world <= SLOT_MAX_W && cap >= 1 && src && recv_idx && local_rows, "slot_ids: bad arguments");
  slot_id_srcs p;
  for (int w = 0; w < SLOT_MAX_W; ++w) {
    p.p[w] = w < world ? src[w] : nullptr;
    TTAM_CHECK_ARG(w >= world || p.p[w], "slot_ids: source %d is null", w);
  }
  const int64_t n = world * cap;
  slot_ids_kernel<<<(int)std::min<int64_t>(ceil_div(n, 256), 1184), 256, 0, (cudaStream_t)stream>>>(p, (int)world, cap, recv_idx,
                                                                                                    local_rows);
  TTAM_LAUNCH_CHECK();
  return TTAM_OK;
}
