// cuTensorMapEncodeTiled through the runtime's driver-entry-point query (no link-time libcuda).
#pragma once
#include "common.cuh"
#include <cuda.h>
#include <mutex>

namespace synt {

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline EncodeTiledFn get_encode_tiled() {
    static EncodeTiledFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* sym = nullptr;
        cudaDriverEntryPointQueryResult qres;
        cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres);
        if (e == cudaSuccess && qres == cudaDriverEntryPointSuccess) fn = reinterpret_cast<EncodeTiledFn>(sym);
    });
    if (!fn) throw Error(-4, "cuTensorMapEncodeTiled not available from the driver");
    return fn;
}

// bf16 tensor, `rank` dims (innermost first), byte strides for dims 1.., SWIZZLE_128B, OOB -> 0
inline void encode_bf16_sw128(CUtensorMap* m, const void* base, int rank, const cuuint64_t* dims,
                              const cuuint64_t* strides_bytes, const cuuint32_t* box, const char* what) {
    cuuint32_t es[5] = {1, 1, 1, 1, 1};
    CUresult r = get_encode_tiled()(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(base), dims,
                                    strides_bytes, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                    CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) throw Error(-4, std::string("cuTensorMapEncodeTiled(") + what + ") failed: " + std::to_string((int)r));
}

}  // namespace synt
