// Host-side TMA tensor-map construction through the driver entry point (no libcuda link dependency).
#pragma once

#include <cuda.h>
#include <cuda_runtime.h>

#include <mutex>

namespace tdb {

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)p;
    });
    return fn;
}

// 2-D bf16 tensor map: dim0 (contiguous) = cols, dim1 = rows with pitch `pitch_elems`; the box is
// box_cols x box_rows with the swizzle span equal to the box row (32/64/128 bytes).
inline bool make_map_2d_bf16(CUtensorMap* m, const void* base, uint64_t cols, uint64_t rows, uint64_t pitch_elems,
                             uint32_t box_cols, uint32_t box_rows) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) return false;
    cuuint64_t dims[2] = {cols, rows};
    cuuint64_t strides[1] = {pitch_elems * 2};
    cuuint32_t box[2] = {box_cols, box_rows};
    cuuint32_t estr[2] = {1, 1};
    const uint32_t row_bytes = box_cols * 2;
    CUtensorMapSwizzle sw = row_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B
                                             : (row_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
    return fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
              CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// General bf16 tensor map of rank <= 5: dims[0] is contiguous; strides_elems[i] is the pitch of dims[i+1]
// in elements (may alias/overlap: used for the overlapping row-group views of the folded convolution).
inline bool make_map_bf16(CUtensorMap* m, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_elems,
                          const uint32_t* box) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) return false;
    cuuint64_t d[5], st[4];
    cuuint32_t bx[5], es[5];
    for (int i = 0; i < rank; ++i) {
        d[i] = dims[i];
        bx[i] = box[i];
        es[i] = 1;
        if (i > 0) st[i - 1] = strides_elems[i - 1] * 2;
    }
    const uint32_t row_bytes = box[0] * 2;
    CUtensorMapSwizzle sw = row_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B
                                             : (row_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
    return fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(base), d, st, bx, es,
              CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

}  // namespace tdb
