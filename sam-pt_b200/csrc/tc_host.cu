// Host-side TMA descriptor construction.  cuTensorMapEncodeTiled is a driver-API symbol; it is resolved at run time
// through cudaGetDriverEntryPoint so the library links against libcudart only (the build box has no libcuda).
#include "common.cuh"
#include "tc_common.cuh"
#include <cudaTypedefs.h>

namespace sampt {

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

int make_tmap_2d_f16(CUtensorMap* out, const void* base, uint64_t inner, uint64_t outer, uint64_t row_stride_bytes,
                     uint32_t box_inner, uint32_t box_outer) {
  EncodeTiledFn enc = get_encode();
  SAMPT_CHECK(enc != nullptr, "cuTensorMapEncodeTiled could not be resolved from the driver");
  SAMPT_CHECK((reinterpret_cast<uintptr_t>(base) & 15) == 0 && (row_stride_bytes & 15) == 0,
              "TMA needs 16-byte aligned base/strides (base=%p stride=%llu)", base, (unsigned long long)row_stride_bytes);
  SAMPT_CHECK(box_inner * 2 == 128 && box_outer <= 256, "TMA box must be 64 halves (128 B swizzle) x <=256 rows");
  cuuint64_t dims[2] = {inner, outer};
  cuuint64_t strides[1] = {row_stride_bytes};
  cuuint32_t box[2] = {box_inner, box_outer};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  SAMPT_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(2d) failed with %d (inner=%llu outer=%llu stride=%llu box=%ux%u)", (int)r,
              (unsigned long long)inner, (unsigned long long)outer, (unsigned long long)row_stride_bytes, box_inner, box_outer);
  return 0;
}

int make_tmap_3d_f16(CUtensorMap* out, const void* base, uint64_t d0, uint64_t d1, uint64_t d2, uint64_t stride1_bytes,
                     uint64_t stride2_bytes, uint32_t box0, uint32_t box1, uint32_t box2) {
  EncodeTiledFn enc = get_encode();
  SAMPT_CHECK(enc != nullptr, "cuTensorMapEncodeTiled could not be resolved from the driver");
  SAMPT_CHECK((reinterpret_cast<uintptr_t>(base) & 15) == 0 && (stride1_bytes & 15) == 0 && (stride2_bytes & 15) == 0,
              "TMA needs 16-byte aligned base/strides");
  SAMPT_CHECK(box0 * 2 == 128 && box1 <= 256 && box2 <= 256, "TMA box must be 64 halves x <=256 x <=256");
  cuuint64_t dims[3] = {d0, d1, d2};
  cuuint64_t strides[2] = {stride1_bytes, stride2_bytes};
  cuuint32_t box[3] = {box0, box1, box2};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  SAMPT_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(3d) failed with %d", (int)r);
  return 0;
}

}  // namespace sampt
