// Uniform negative sampling with rejection of a user's known positives (reference src/data/samplers.py:11-85).
// The reference loops over the batch in Python (one randint + torch.isin round trip per row, ~30 us/row); here one
// thread draws one (row, slot) and re-draws it for as long as it hits a positive, at most 1 + max_rounds draws
// (the reference gives up after 10 re-sampling rounds with a RuntimeError: here a device flag is raised and the caller
// throws).  Positives are a sorted int64 key array user * num_items + item, searched by bisection.
// Statistical (not bit) parity: the draws come from Philox4x32-10 keyed by (seed, offset + element, round).
#include "common.cuh"

namespace ttam {

__device__ __forceinline__ bool key_present(const int64_t* __restrict__ keys, int64_t n, int64_t key) {
  int64_t lo = 0, hi = n;
  while (lo < hi) {
    const int64_t mid = (lo + hi) >> 1;
    const int64_t v = keys[mid];
    if (v < key) lo = mid + 1;
    else hi = mid;
  }
  return lo < n && keys[lo] == key;
}

__global__ void __launch_bounds__(256) sample_negatives_kernel(const int64_t* __restrict__ users, int64_t B, int64_t N,
                                                               int64_t num_items, const int64_t* __restrict__ pos_keys,
                                                               int64_t n_keys, int max_rounds, uint64_t seed, uint64_t offset,
                                                               const ttam_step_state* __restrict__ st,
                                                               int64_t* __restrict__ out, int32_t* __restrict__ fail) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= B * N) return;
  const int64_t u = users[e / N];
  const uint64_t base = offset + (st ? st->rng_offset : 0ull) + (uint64_t)e * (uint64_t)(max_rounds + 1);
  int64_t item = 0;
  bool ok = false;
  for (int r = 0; r <= max_rounds && !ok; ++r) {
    const uint4 rnd = philox4x32(seed, base + (uint64_t)r);
    const uint64_t r64 = ((uint64_t)rnd.x << 32) | (uint64_t)rnd.y;
    item = (int64_t)__umul64hi(r64, (uint64_t)num_items);  // uniform on [0, num_items)
    ok = n_keys == 0 || !key_present(pos_keys, n_keys, u * num_items + item);
  }
  out[e] = item;
  if (!ok) atomicExch(fail, 1);
}

}  // namespace ttam

using namespace ttam;

extern "C" int ttam_sample_negatives(const int64_t* users, int64_t B, int64_t N, int64_t num_items, const int64_t* pos_keys,
                                     int64_t n_keys, int max_rounds, uint64_t seed, uint64_t offset,
                                     const ttam_step_state* state_dev, int64_t* out, int32_t* fail_flag, void* stream) {
  TTAM_CHECK_ARG(N > 0, "num_negatives must be greater than zero.");
  TTAM_CHECK_ARG(num_items > 1, "num_items must be greater than one.");
  TTAM_CHECK_ARG(B >= 0 && max_rounds >= 0 && n_keys >= 0, "sample_negatives: bad size");
  if (B == 0) return TTAM_OK;
  TTAM_CHECK_ARG(users && out && fail_flag && (n_keys == 0 || pos_keys), "sample_negatives: null pointer");
  const int64_t total = B * N;
  sample_negatives_kernel<<<(unsigned)ceil_div(total, 256), 256, 0, (cudaStream_t)stream>>>(
      users, B, N, num_items, pos_keys, n_keys, max_rounds, seed, offset, state_dev, out, fail_flag);
  TTAM_LAUNCH_CHECK();
  return TTAM_OK;
}
