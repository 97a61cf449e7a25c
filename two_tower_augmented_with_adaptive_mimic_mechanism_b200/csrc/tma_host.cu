// Host side of TMA: cuTensorMapEncodeTiled through the runtime's driver entry point, so that libttam.so does
// not need libcuda at link time (the library must load on a CPU-only build box to be symbol-checked).
#include "common.cuh"
#include "sm100.cuh"

namespace ttam {

PFN_encodeTiled get_encode_tiled() {
  static PFN_encodeTiled fn = nullptr;
  if (fn) return fn;
  void* p = nullptr;
  cudaDriverEntryPointQueryResult qres;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) != cudaSuccess ||
      qres != cudaDriverEntryPointSuccess || p == nullptr) {
    cudaGetLastError();
    return nullptr;
  }
  fn = reinterpret_cast<PFN_encodeTiled>(p);
  return fn;
}

int make_tmap_2d(CUtensorMap* out, const void* base, CUtensorMapDataType dtype, int elem_bytes, uint64_t rows,
                 uint64_t cols, uint64_t row_stride_bytes, uint32_t box_cols, uint32_t box_rows,
                 CUtensorMapSwizzle swizzle) {
  PFN_encodeTiled enc = get_encode_tiled();
  if (!enc) {
    set_error("cuTensorMapEncodeTiled is not available from this driver");
    return TTAM_ECUDA;
  }
  (void)elem_bytes;
  cuuint64_t gdim[2] = {cols, rows};
  cuuint64_t gstride[1] = {row_stride_bytes};
  cuuint32_t box[2] = {box_cols, box_rows};
  cuuint32_t estride[2] = {1, 1};
  CUresult r = enc(out, dtype, 2, const_cast<void*>(base), gdim, gstride, box, estride, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   swizzle, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed with CUresult %d (rows=%llu cols=%llu stride=%llu box=%ux%u)", (int)r,
              (unsigned long long)rows, (unsigned long long)cols, (unsigned long long)row_stride_bytes, box_cols, box_rows);
    return TTAM_ECUDA;
  }
  return TTAM_OK;
}

}  // namespace ttam
