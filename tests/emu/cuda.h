// Stand-in for <cuda.h> in the host emulation build (tests/emu): the tensor-map descriptor and the
// encoder the library fetches through cudaGetDriverEntryPoint.  Test infrastructure.
#pragma once
#include <cstdint>
typedef unsigned int cuuint32_t;
typedef unsigned long long cuuint64_t;
typedef int CUresult;
enum { CUDA_SUCCESS = 0, CUDA_ERROR_INVALID_VALUE = 1 };
enum CUtensorMapDataType { CU_TENSOR_MAP_DATA_TYPE_FLOAT64 = 10 };
enum CUtensorMapInterleave { CU_TENSOR_MAP_INTERLEAVE_NONE = 0 };
enum CUtensorMapSwizzle { CU_TENSOR_MAP_SWIZZLE_NONE = 0, CU_TENSOR_MAP_SWIZZLE_32B, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_SWIZZLE_128B };
enum CUtensorMapL2promotion { CU_TENSOR_MAP_L2_PROMOTION_NONE = 0 };
enum CUtensorMapFloatOOBfill { CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE = 0 };
// the emulated descriptor: a rank-2 tile of 8-byte elements
struct alignas(64) CUtensorMap {
    unsigned char* base;
    unsigned long long dim0, dim1, stride1_bytes;
    unsigned box0, box1, swizzle_mask;
    unsigned long long pad[10];
};
inline CUresult emu_cuTensorMapEncodeTiled(CUtensorMap* tm, CUtensorMapDataType dt, cuuint32_t rank, void* base, const cuuint64_t* dims,
                                           const cuuint64_t* strides, const cuuint32_t* box, const cuuint32_t*, CUtensorMapInterleave,
                                           CUtensorMapSwizzle sw, CUtensorMapL2promotion, CUtensorMapFloatOOBfill) {
    if (dt != CU_TENSOR_MAP_DATA_TYPE_FLOAT64 || rank != 2) return CUDA_ERROR_INVALID_VALUE;
    if (((uintptr_t)base & 15) || (strides[0] & 15)) return CUDA_ERROR_INVALID_VALUE;          // TMA alignment rules
    const unsigned row_bytes = box[0] * 8;
    if (sw == CU_TENSOR_MAP_SWIZZLE_32B && row_bytes > 32) return CUDA_ERROR_INVALID_VALUE;   // inner box within the swizzle span
    if (sw == CU_TENSOR_MAP_SWIZZLE_64B && row_bytes > 64) return CUDA_ERROR_INVALID_VALUE;
    if (sw == CU_TENSOR_MAP_SWIZZLE_128B && row_bytes > 128) return CUDA_ERROR_INVALID_VALUE;
    if (box[0] > 256 || box[1] > 256 || (row_bytes & 15)) return CUDA_ERROR_INVALID_VALUE;
    tm->base = (unsigned char*)base; tm->dim0 = dims[0]; tm->dim1 = dims[1]; tm->stride1_bytes = strides[0];
    tm->box0 = box[0]; tm->box1 = box[1];
    tm->swizzle_mask = sw == CU_TENSOR_MAP_SWIZZLE_128B ? 7 : sw == CU_TENSOR_MAP_SWIZZLE_64B ? 3 : sw == CU_TENSOR_MAP_SWIZZLE_32B ? 1 : 0;
    return CUDA_SUCCESS;
}
