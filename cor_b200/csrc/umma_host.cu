// Host side of the TMA path: encode CUtensorMap through the driver entry point obtained at run time
// (cudaGetDriverEntryPoint), so that the library links only the static CUDA runtime and still loads on
// a machine without a driver.
#include "umma.cuh"

namespace cor {
namespace umma {

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (fn) return fn;
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess || !p) {
    set_error("cuTensorMapEncodeTiled not available from the driver");
    return nullptr;
  }
  fn = reinterpret_cast<EncodeTiledFn>(p);
  return fn;
}

// The driver call below needs a CUDA context current on the calling thread.  A thread that has only ever been handed
// tensors (a fresh Python thread, autograd's backward thread when our node runs first) may have none yet, and the
// runtime binds one lazily only on ITS OWN calls: bind the primary context of the device that owns `ptr`.
static void bind_context_of(const void* ptr) {
  typedef CUresult (*PtrAttrFn)(void*, CUpointer_attribute, CUdeviceptr);
  static PtrAttrFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuPointerGetAttribute", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PtrAttrFn>(p);
  }
  int ordinal = -1;
  if (!fn || fn(&ordinal, CU_POINTER_ATTRIBUTE_DEVICE_ORDINAL, (CUdeviceptr)(uintptr_t)ptr) != CUDA_SUCCESS || ordinal < 0) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, ptr) == cudaSuccess && a.type == cudaMemoryTypeDevice) ordinal = a.device;
    else if (cudaGetDevice(&ordinal) != cudaSuccess) ordinal = 0;
    cudaGetLastError();
  }
  cudaSetDevice(ordinal);
}

int encode_tmap_bf16_2d(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint32_t box_rows, uint32_t box_cols) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return COR_ECUDA;
  if (((uintptr_t)base & 15) != 0 || (cols * 2) % 16 != 0) {
    set_error("TMA operand must be 16-byte aligned with a 16-byte multiple row pitch (cols=%llu)", (unsigned long long)cols);
    return COR_EINVAL;
  }
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {cols * 2};
  cuuint32_t box[2] = {box_cols, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = CUDA_SUCCESS;
  for (int attempt = 0; attempt < 2; ++attempt) {
    r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_ERROR_INVALID_CONTEXT && r != CUDA_ERROR_NOT_INITIALIZED) break;
    bind_context_of(base);            // no context on this thread yet: bind the owner's and try once more
  }
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed with CUresult %d (rows=%llu cols=%llu box=%ux%u)", (int)r, (unsigned long long)rows,
              (unsigned long long)cols, box_rows, box_cols);
    return COR_ECUDA;
  }
  return COR_OK;
}

int encode_tmap_tiled(CUtensorMap* out, const void* base, int elem_bytes, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                      const uint32_t* box) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return COR_ECUDA;
  if (rank < 1 || rank > 5 || ((uintptr_t)base & 15) != 0) {
    set_error("encode_tmap_tiled: rank %d / base alignment", rank);
    return COR_EINVAL;
  }
  CUtensorMapDataType dt = elem_bytes == 4 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : elem_bytes == 2 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16
                                                                                               : CU_TENSOR_MAP_DATA_TYPE_UINT8;
  cuuint64_t d[5], st[4];
  cuuint32_t bx[5], es[5];
  for (int i = 0; i < rank; ++i) { d[i] = dims[i]; bx[i] = box[i]; es[i] = 1; }
  for (int i = 0; i + 1 < rank; ++i) {
    st[i] = strides_bytes[i];
    if (st[i] % 16 != 0) {
      set_error("encode_tmap_tiled: stride %d = %llu bytes is not a multiple of 16", i, (unsigned long long)st[i]);
      return COR_EINVAL;
    }
  }
  if (((uint64_t)bx[0] * elem_bytes) % 16 != 0) {
    set_error("encode_tmap_tiled: inner box of %u x %d bytes is not a multiple of 16", bx[0], elem_bytes);
    return COR_EINVAL;
  }
  CUresult r = CUDA_SUCCESS;
  for (int attempt = 0; attempt < 2; ++attempt) {
    r = enc(out, dt, (cuuint32_t)rank, const_cast<void*>(base), d, st, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
            CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_ERROR_INVALID_CONTEXT && r != CUDA_ERROR_NOT_INITIALIZED) break;
    bind_context_of(base);
  }
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled (rank %d, elem %d B) failed with CUresult %d", rank, elem_bytes, (int)r);
    return COR_ECUDA;
  }
  return COR_OK;
}

}  // namespace umma
}  // namespace cor
