// Category-alignment loss (reference src/pipelines/training.py:530-579, applied to the step's item embeddings at
// :805-820) and its gradient, entirely on the device: no host synchronisation, CUDA-graph capturable.
//
//   L_cal = mean over the non-major categories c with >= 2 rows of || Cov(rows of c) - Cov(rows of the major category) ||_F^2
//   (unbiased covariances of the D-dim embeddings whose primary category is c; zero when the major category has < 2 rows,
//   when only one category is present or when nothing can be compared).
//
// The reference loops over the categories in Python (and halves its throughput when this term is on).  Here:
//   1. keys = cat[item_idx]; stable radix sort (ttam_sort_rows) -> rows grouped by category
//   2. segment bounds by bisection; every category is cut into chunks of 128 rows (one block each, fixed order:
//      deterministic), chunk -> category map by a one-block scan
//   3. per chunk: partial sums -> means; per chunk: partial D x D covariance of the centred rows -> covariances
//   4. per category: diff = Cov_c - Cov_major, sum of squares; one block combines them into the loss, the number of
//      compared categories n_c and G_major = -(2/n_c) sum_c diff_c;  G_c = (2/n_c) diff_c
//   5. per chunk: grad rows  2/(n-1) (x - mean_c) G_c, added with weight lambda_c into dL/do_i (and dL/dq of the positives)
#include "common.cuh"

namespace ttam {
namespace cal {

constexpr int kChunk = 128;   // rows per block
constexpr int kThreads = 256;
constexpr int kMaxD = 128;

struct P {
  const int64_t* item_idx;
  const int64_t* cat;
  const float* emb;
  int64_t R, B, num_items;
  int D, n_cat, major;
  float lambda_c;
  // workspace
  int64_t* keys;
  int64_t* sorted;
  int32_t* perm;
  int32_t* start;       // [n_cat + 1] first sorted position of category c
  int32_t* chunk_base;  // [n_cat + 1] first chunk of category c
  float* psum;          // [max_chunks][D]
  float* mean;          // [n_cat][D]
  float* pcov;          // [max_chunks][D*D]
  float* cov;           // [n_cat][D*D]  (later: diff_c)
  float* ssq;           // [n_cat]
  float* gmajor;        // [D*D]
  float* scal;          // [0] = 2 / n_c (0: nothing to do), [1] = loss
  int max_chunks;
  // outputs
  float* loss_out;
  float* cal_out;
  float* grad_a;
  float* grad_b;
};

__global__ void keys_kernel(P p) {
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= p.R) return;
  const int64_t it = p.item_idx[r];
  int64_t c = (it >= 0 && it < p.num_items) ? p.cat[it] : 0;
  if (c < 0 || c >= p.n_cat) c = 0;
  p.keys[r] = c;
}

__global__ void bounds_kernel(P p) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c > p.n_cat) return;
  int64_t lo = 0, hi = p.R;
  while (lo < hi) {
    const int64_t mid = (lo + hi) >> 1;
    if (p.sorted[mid] < (int64_t)c) lo = mid + 1;
    else hi = mid;
  }
  p.start[c] = (int32_t)lo;
}

// exclusive scan of the per-category chunk counts (one block; n_cat is a few hundred to a few thousand)
__global__ void __launch_bounds__(1024) chunk_scan_kernel(P p) {
  __shared__ int32_t carry;
  __shared__ int32_t buf[1024];
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (int c0 = 0; c0 <= p.n_cat; c0 += 1024) {
    const int c = c0 + threadIdx.x;
    int32_t n = 0;
    if (c < p.n_cat) n = (p.start[c + 1] - p.start[c] + kChunk - 1) / kChunk;
    buf[threadIdx.x] = n;
    __syncthreads();
    for (int off = 1; off < 1024; off <<= 1) {  // Hillis-Steele inclusive scan
      int32_t v = threadIdx.x >= off ? buf[threadIdx.x - off] : 0;
      __syncthreads();
      buf[threadIdx.x] += v;
      __syncthreads();
    }
    if (c <= p.n_cat) p.chunk_base[c] = carry + buf[threadIdx.x] - n;
    __syncthreads();
    if (threadIdx.x == 1023) carry += buf[1023];
    __syncthreads();
  }
}

// which category does chunk `ch` belong to, and which rows does it cover?  (-1: no such chunk)
__device__ __forceinline__ int chunk_category(const P& p, int ch, int& row_lo, int& row_hi) {
  if (ch >= p.chunk_base[p.n_cat]) return -1;
  int lo = 0, hi = p.n_cat;  // last c with chunk_base[c] <= ch
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (p.chunk_base[mid] <= ch) lo = mid;
    else hi = mid;
  }
  // categories without rows have chunk_base[c] == chunk_base[c+1]: `lo` is the last one at or before ch, which has rows
  const int c = lo;
  row_lo = p.start[c] + (ch - p.chunk_base[c]) * kChunk;
  row_hi = min(p.start[c + 1], row_lo + kChunk);
  return c;
}

__global__ void __launch_bounds__(kThreads) mean_partial_kernel(P p) {
  int lo, hi;
  const int c = chunk_category(p, blockIdx.x, lo, hi);
  if (c < 0) return;
  for (int d = threadIdx.x; d < p.D; d += blockDim.x) {
    float s = 0.f;
    for (int j = lo; j < hi; ++j) s += p.emb[(int64_t)p.perm[j] * p.D + d];
    p.psum[(int64_t)blockIdx.x * p.D + d] = s;
  }
}

__global__ void mean_final_kernel(P p) {
  const int c = blockIdx.x;
  const int n = p.start[c + 1] - p.start[c];
  for (int d = threadIdx.x; d < p.D; d += blockDim.x) {
    float s = 0.f;
    for (int ch = p.chunk_base[c]; ch < p.chunk_base[c + 1]; ++ch) s += p.psum[(int64_t)ch * p.D + d];
    p.mean[(int64_t)c * p.D + d] = n > 0 ? s / (float)n : 0.f;
  }
}

// partial covariance of one chunk: sum over its rows of (x - mean)(x - mean)^T.  Thread t owns the entries
// t, t + 256, ... of the D x D matrix; the centred row is broadcast from shared memory.
__global__ void __launch_bounds__(kThreads) cov_partial_kernel(P p) {
  __shared__ float cen[kMaxD];
  int lo, hi;
  const int c = chunk_category(p, blockIdx.x, lo, hi);
  if (c < 0) return;
  const int D = p.D, DD = D * D;
  constexpr int kPer = (kMaxD * kMaxD + kThreads - 1) / kThreads;  // 64 entries per thread at D = 128
  float acc[kPer];
#pragma unroll
  for (int i = 0; i < kPer; ++i) acc[i] = 0.f;
  for (int j = lo; j < hi; ++j) {
    __syncthreads();
    for (int d = threadIdx.x; d < D; d += blockDim.x) cen[d] = p.emb[(int64_t)p.perm[j] * D + d] - p.mean[(int64_t)c * D + d];
    __syncthreads();
#pragma unroll
    for (int i = 0; i < kPer; ++i) {
      const int e = threadIdx.x + i * kThreads;
      if (e < DD) acc[i] = fmaf(cen[e / D], cen[e % D], acc[i]);
    }
  }
#pragma unroll
  for (int i = 0; i < kPer; ++i) {
    const int e = threadIdx.x + i * kThreads;
    if (e < DD) p.pcov[(int64_t)blockIdx.x * DD + e] = acc[i];
  }
}

__global__ void __launch_bounds__(kThreads) cov_final_kernel(P p) {
  const int c = blockIdx.y;
  const int DD = p.D * p.D;
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= DD) return;
  const int n = p.start[c + 1] - p.start[c];
  float s = 0.f;
  for (int ch = p.chunk_base[c]; ch < p.chunk_base[c + 1]; ++ch) s += p.pcov[(int64_t)ch * DD + e];
  p.cov[(int64_t)c * DD + e] = n > 1 ? s / (float)(n - 1) : 0.f;  // _compute_covariance: zeros for <= 1 row
}

__device__ __forceinline__ bool compared(const P& p, int c) {
  return c != p.major && (p.start[c + 1] - p.start[c]) >= 2;
}

// diff_c = Cov_c - Cov_major (stored over Cov_c), ssq[c] = sum diff_c^2   (block per category, fixed-order tree)
__global__ void __launch_bounds__(kThreads) diff_kernel(P p) {
  __shared__ float red[kThreads];
  const int c = blockIdx.x;
  const int DD = p.D * p.D;
  float s = 0.f;
  if (compared(p, c)) {
    for (int e = threadIdx.x; e < DD; e += blockDim.x) {
      const float d = p.cov[(int64_t)c * DD + e] - p.cov[(int64_t)p.major * DD + e];
      p.cov[(int64_t)c * DD + e] = d;  // only this block touches category c; the major's matrix stays intact
      s = fmaf(d, d, s);
    }
  }
  red[threadIdx.x] = s;
  __syncthreads();
  for (int off = kThreads / 2; off > 0; off >>= 1) {
    if (threadIdx.x < off) red[threadIdx.x] += red[threadIdx.x + off];
    __syncthreads();
  }
  if (threadIdx.x == 0) p.ssq[c] = red[0];
}

// loss, number of compared categories, scale 2 / n_c; adds lambda_c * loss into loss_out[0]   (one block)
__global__ void __launch_bounds__(kThreads) finish_kernel(P p) {
  __shared__ float red[kThreads];
  __shared__ int cnt[kThreads];
  const int n_major = p.start[p.major + 1] - p.start[p.major];
  float s = 0.f;
  int n = 0;
  for (int c = threadIdx.x; c < p.n_cat; c += blockDim.x) {
    if (compared(p, c)) {
      s += p.ssq[c];
      ++n;
    }
  }
  red[threadIdx.x] = s;
  cnt[threadIdx.x] = n;
  __syncthreads();
  for (int off = kThreads / 2; off > 0; off >>= 1) {
    if (threadIdx.x < off) {
      red[threadIdx.x] += red[threadIdx.x + off];
      cnt[threadIdx.x] += cnt[threadIdx.x + off];
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    const int n_c = n_major >= 2 ? cnt[0] : 0;  // training.py:559-560: nothing without two rows of the major category
    const float loss = n_c > 0 ? red[0] / (float)n_c : 0.f;
    p.scal[0] = n_c > 0 ? 2.f / (float)n_c : 0.f;
    p.scal[1] = loss;
    if (p.cal_out) p.cal_out[0] = loss;
    if (p.loss_out) p.loss_out[0] += p.lambda_c * loss;
  }
}

// G_major[e] = -(2/n_c) sum over compared c of diff_c[e]   (fixed order over c)
__global__ void __launch_bounds__(kThreads) gmajor_kernel(P p) {
  const int DD = p.D * p.D;
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= DD) return;
  const float sc = p.scal[0];
  float s = 0.f;
  if (sc != 0.f)
    for (int c = 0; c < p.n_cat; ++c)
      if (compared(p, c)) s += p.cov[(int64_t)c * DD + e];
  p.gmajor[e] = -sc * s;
}

// grad rows of one chunk: y = 2/(n-1) (x - mean_c) G_c, with G_c = (2/n_c) diff_c or G_major; G_c staged in shared memory
__global__ void __launch_bounds__(kThreads) grad_kernel(P p) {
  extern __shared__ float sm[];  // G [D*D] | cen [2][D]
  int lo, hi;
  const int c = chunk_category(p, blockIdx.x, lo, hi);
  if (c < 0) return;
  const float sc = p.scal[0];
  const int n = p.start[c + 1] - p.start[c];
  const bool is_major = c == p.major;
  if (sc == 0.f || n < 2 || !(is_major || compared(p, c))) return;
  const int D = p.D, DD = D * D;
  float* G = sm;
  float* cen = sm + DD;
  for (int e = threadIdx.x; e < DD; e += blockDim.x) G[e] = is_major ? p.gmajor[e] : sc * p.cov[(int64_t)c * DD + e];
  const float w = p.lambda_c * 2.f / (float)(n - 1);
  // two rows per pass: threads [0, 128) serve row j, threads [128, 256) row j + 1; each thread one output column (D <= 128)
  const int half = threadIdx.x >> 7, col = threadIdx.x & 127;
  for (int j0 = lo; j0 < hi; j0 += 2) {
    const int j = j0 + half;
    __syncthreads();
    int64_t row = -1;
    if (j < hi) {
      row = p.perm[j];
      if (col < D) cen[half * kMaxD + col] = p.emb[row * D + col] - p.mean[(int64_t)c * D + col];
    }
    __syncthreads();
    if (j < hi && col < D) {
      float y = 0.f;
      const float* cr = cen + half * kMaxD;
      for (int d = 0; d < D; ++d) y = fmaf(cr[d], G[d * D + col], y);
      y *= w;
      p.grad_a[row * D + col] += y;
      if (p.grad_b && row < p.B) p.grad_b[row * D + col] += y;
    }
  }
}

struct Ws {
  int64_t off = 0;
  char* base;
  explicit Ws(void* b) : base((char*)b) {}
  template <class T>
  T* take(int64_t n) {
    T* ptr = base ? (T*)(base + off) : nullptr;
    off += align_up(n * (int64_t)sizeof(T), 256);
    return ptr;
  }
};

static int64_t carve(P& p, void* base, int64_t sort_ws_bytes, void** sort_ws) {
  Ws w(base);
  const int64_t DD = (int64_t)p.D * p.D;
  p.max_chunks = (int)(p.R / kChunk + p.n_cat + 1);
  p.keys = w.take<int64_t>(p.R);
  p.sorted = w.take<int64_t>(p.R);
  p.perm = w.take<int32_t>(p.R);
  p.start = w.take<int32_t>(p.n_cat + 2);
  p.chunk_base = w.take<int32_t>(p.n_cat + 2);
  p.psum = w.take<float>((int64_t)p.max_chunks * p.D);
  p.mean = w.take<float>((int64_t)p.n_cat * p.D);
  p.pcov = w.take<float>((int64_t)p.max_chunks * DD);
  p.cov = w.take<float>((int64_t)p.n_cat * DD);
  p.ssq = w.take<float>(p.n_cat);
  p.gmajor = w.take<float>(DD);
  p.scal = w.take<float>(64);
  char* s = w.take<char>(sort_ws_bytes);
  if (sort_ws) *sort_ws = s;
  return w.off + 256;
}

}  // namespace cal
}  // namespace ttam

using namespace ttam;
using namespace ttam::cal;

extern "C" int64_t ttam_category_alignment_workspace_bytes(int64_t R, int64_t D, int64_t n_categories) {
  if (R <= 0 || D <= 0 || n_categories <= 0) return 256;
  P p{};
  p.R = R; p.D = (int)D; p.n_cat = (int)n_categories;
  return carve(p, nullptr, ttam_sort_workspace_bytes(R), nullptr);
}

extern "C" int ttam_category_alignment(const int64_t* item_idx, int64_t R, const float* emb, int64_t D,
                                       const int64_t* cat_tensor, int64_t num_items, int64_t n_categories, int64_t major,
                                       float lambda_c, float* loss_out, float* cal_out, float* grad_a, float* grad_b,
                                       int64_t B, void* workspace, int64_t workspace_bytes, void* stream) {
  TTAM_CHECK_ARG(R >= 0 && D > 0 && n_categories > 0 && major >= 0 && major < n_categories && num_items > 0,
                 "category_alignment: bad argument");
  if (D > kMaxD) {
    set_error("category_alignment: D = %lld > %d is not supported by this kernel", (long long)D, kMaxD);
    return TTAM_EUNSUPPORTED;
  }
  TTAM_CHECK_ARG(n_categories < (1 << 20) && R < (1ll << 31), "category_alignment: too many categories / rows");
  if (R == 0) return TTAM_OK;
  TTAM_CHECK_ARG(item_idx && emb && cat_tensor && workspace && grad_a, "category_alignment: null pointer");
  P p{};
  p.item_idx = item_idx; p.cat = cat_tensor; p.emb = emb; p.R = R; p.B = B; p.num_items = num_items;
  p.D = (int)D; p.n_cat = (int)n_categories; p.major = (int)major; p.lambda_c = lambda_c;
  p.loss_out = loss_out; p.cal_out = cal_out; p.grad_a = grad_a; p.grad_b = grad_b;
  const int64_t sort_ws = ttam_sort_workspace_bytes(R);
  void* sws = nullptr;
  const int64_t need = carve(p, workspace, sort_ws, &sws);
  if (workspace_bytes < need) {
    set_error("category_alignment: workspace too small (%lld < %lld)", (long long)workspace_bytes, (long long)need);
    return TTAM_EWORKSPACE;
  }
  cudaStream_t st = (cudaStream_t)stream;
  const int DD = (int)(D * D);
  keys_kernel<<<(unsigned)ceil_div(R, 256), 256, 0, st>>>(p);
  TTAM_LAUNCH_CHECK();
  const int rc = ttam_sort_rows(p.keys, R, n_categories, p.sorted, p.perm, sws, sort_ws, stream);
  if (rc != TTAM_OK) return rc;
  bounds_kernel<<<(unsigned)ceil_div(n_categories + 1, 256), 256, 0, st>>>(p);
  TTAM_LAUNCH_CHECK();
  chunk_scan_kernel<<<1, 1024, 0, st>>>(p);
  TTAM_LAUNCH_CHECK();
  mean_partial_kernel<<<p.max_chunks, kThreads, 0, st>>>(p);
  TTAM_LAUNCH_CHECK();
  mean_final_kernel<<<(unsigned)n_categories, 128, 0, st>>>(p);
  TTAM_LAUNCH_CHECK();
  cov_partial_kernel<<<p.max_chunks, kThreads, 0, st>>>(p);
  TTAM_LAUNCH_CHECK();
  cov_final_kernel<<<dim3((unsigned)ceil_div(DD, kThreads), (unsigned)n_categories), kThreads, 0, st>>>(p);
  TTAM_LAUNCH_CHECK();
  diff_kernel<<<(unsigned)n_categories, kThreads, 0, st>>>(p);
  TTAM_LAUNCH_CHECK();
  finish_kernel<<<1, kThreads, 0, st>>>(p);
  TTAM_LAUNCH_CHECK();
  gmajor_kernel<<<(unsigned)ceil_div(DD, kThreads), kThreads, 0, st>>>(p);
  TTAM_LAUNCH_CHECK();
  const size_t smem = (size_t)(DD + 2 * kMaxD) * sizeof(float);
  static bool attr_done = false;
  if (!attr_done) {
    TTAM_CUDA(cudaFuncSetAttribute(grad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)((kMaxD * kMaxD + 2 * kMaxD) * sizeof(float))));
    attr_done = true;
  }
  grad_kernel<<<p.max_chunks, kThreads, smem, st>>>(p);
  TTAM_LAUNCH_CHECK();
  return TTAM_OK;
}
