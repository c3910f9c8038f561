// TMA tensor-map construction. cuTensorMapEncodeTiled is fetched through the runtime's
// driver-entry-point query so the library links against cudart only (no libcuda stub).
#include "conv_tc.h"

#include <mutex>

namespace cgb {

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres);
    if (e == cudaSuccess && qres == cudaDriverEntryPointSuccess) fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  CGB_CHECK(fn != nullptr, "cuTensorMapEncodeTiled unavailable (no CUDA driver?)");
  return fn;
}

static CUtensorMapSwizzle swz(int bytes) {
  switch (bytes) {
    case 128: return CU_TENSOR_MAP_SWIZZLE_128B;
    case 64: return CU_TENSOR_MAP_SWIZZLE_64B;
    case 32: return CU_TENSOR_MAP_SWIZZLE_32B;
    default: break;
  }
  CGB_CHECK(false, "unsupported swizzle span");
  return CU_TENSOR_MAP_SWIZZLE_NONE;
}

CUtensorMap make_tmap_act5d(const bf16* base, const int dims[5], const long long strides_elems[4], int box_c,
                            int box_w, int box_h, int swizzle_bytes) {
  CGB_CHECK(box_c * 2 == swizzle_bytes, "box_c must span exactly one swizzle row");
  CGB_CHECK((reinterpret_cast<uintptr_t>(base) & 15) == 0, "activation base must be 16-byte aligned");
  CUtensorMap m;
  cuuint64_t gdim[5], gstr[4];
  cuuint32_t box[5] = {(cuuint32_t)box_c, (cuuint32_t)box_w, 1u, (cuuint32_t)box_h, 1u};
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  for (int i = 0; i < 5; ++i) {
    CGB_CHECK(dims[i] > 0, "tensor-map dim must be positive");
    gdim[i] = (cuuint64_t)dims[i];
  }
  for (int i = 0; i < 4; ++i) {
    gstr[i] = (cuuint64_t)strides_elems[i] * 2;
    CGB_CHECK(gstr[i] % 16 == 0, "tensor-map stride must be a multiple of 16 bytes");
  }
  CGB_CHECK(box_w <= 256 && box_h <= 256, "TMA box dims are limited to 256");
  CUresult r = get_encode()(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<bf16*>(base), gdim, gstr, box,
                            estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swz(swizzle_bytes),
                            CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  CGB_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(5d) failed with code " + std::to_string((int)r));
  return m;
}

CUtensorMap make_tmap_2d(const bf16* base, long long rows, long long cols, long long pitch_elems, int box_cols,
                         int box_rows, int swizzle_bytes) {
  CGB_CHECK(box_cols * 2 == swizzle_bytes, "box_cols must span exactly one swizzle row");
  CGB_CHECK((reinterpret_cast<uintptr_t>(base) & 15) == 0, "matrix base must be 16-byte aligned");
  CUtensorMap m;
  cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t gstr[1] = {(cuuint64_t)pitch_elems * 2};
  cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CGB_CHECK(gstr[0] % 16 == 0, "matrix pitch must be a multiple of 16 bytes");
  CGB_CHECK(box_rows <= 256, "TMA box dims are limited to 256");
  CUresult r = get_encode()(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<bf16*>(base), gdim, gstr, box, estr,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, swz(swizzle_bytes), CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  CGB_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(2d) failed with code " + std::to_string((int)r));
  return m;
}

CUtensorMap make_tmap_tapn_weights(const bf16* base, int k, long long pitch_elems) {
  CGB_CHECK((reinterpret_cast<uintptr_t>(base) & 15) == 0, "matrix base must be 16-byte aligned");
  CGB_CHECK(k >= 2 && k <= 8, "taps-in-N weights: 2 <= k <= 8");
  CUtensorMap m;
  // element (ci, row, tx, ty) = W[row][(ty * k + tx) * 64 + ci]; tx runs to 8: for k < 8 the extra taps read the next
  // filter row's first taps (in bounds: the matrix has >= 5 rows); their output columns are never gathered
  cuuint64_t gdim[4] = {64, 4, 8, (cuuint64_t)k};
  cuuint64_t gstr[3] = {(cuuint64_t)pitch_elems * 2, 128, (cuuint64_t)k * 128};
  cuuint32_t box[4] = {64, 4, 8, (cuuint32_t)k};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CGB_CHECK(gstr[0] % 16 == 0, "matrix pitch must be a multiple of 16 bytes");
  CUresult r = get_encode()(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<bf16*>(base), gdim, gstr, box, estr,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  CGB_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(4d) failed with code " + std::to_string((int)r));
  return m;
}

}  // namespace cgb
