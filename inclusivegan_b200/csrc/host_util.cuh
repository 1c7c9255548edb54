// Host-side plumbing of libb200knn: error reporting, TMA descriptor encoding, grow-only device buffers, the
// process-wide pinned upload ring.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <thread>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <mutex>
#include <new>
#include <string>
#include <utility>
#include <vector>

#include "../../include/b200knn.h"
#include "kernels.cuh"

namespace {

using namespace b200;

thread_local std::string g_last_error;

int fail(int code, const char *fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_last_error = buf;
    return code;
}

#define CU_TRY(expr)                                                                                    \
    do {                                                                                                \
        cudaError_t _e = (expr);                                                                        \
        if (_e != cudaSuccess)                                                                          \
            return fail(_e == cudaErrorMemoryAllocation ? B200KNN_ENOMEM : B200KNN_ECUDA, "%s failed: %s (%s:%d)", #expr, \
                        cudaGetErrorString(_e), __FILE__, __LINE__);                                    \
    } while (0)
#define TRY(expr)                \
    do {                         \
        int _r = (expr);         \
        if (_r != B200KNN_OK) return _r; \
    } while (0)

// --------------------------------------------------------------------------------------------
// driver entry point for TMA descriptors (resolved at run time: the .so loads without libcuda)
// --------------------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                    const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
PFN_encodeTiled g_encode = nullptr;

int resolve_driver() {
    if (g_encode) return B200KNN_OK;
    void *fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
    if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || !fn)
        return fail(B200KNN_ENODEVICE, "cuTensorMapEncodeTiled not available from the CUDA driver (%s)", cudaGetErrorString(e));
    g_encode = reinterpret_cast<PFN_encodeTiled>(fn);
    return B200KNN_OK;
}

// BF16 row-major [rows, kp] -> 2-D tensor map with a (BK x box_rows) SWIZZLE_128B box
// (f32 = true: fp32 containers of TF32 values, 32 elements per 128-byte box row)
int make_tmap(CUtensorMap *m, const void *base, uint64_t rows, uint64_t kp, uint32_t box_rows, bool f32 = false) {
    TRY(resolve_driver());
    cuuint64_t dims[2] = {kp, rows};
    cuuint64_t strides[1] = {kp * (f32 ? 4u : 2u)};
    cuuint32_t box[2] = {static_cast<cuuint32_t>(f32 ? BK / 2 : BK), box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = g_encode(m, f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void *>(base), dims, strides, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(B200KNN_ECUDA, "cuTensorMapEncodeTiled failed with CUresult %d (rows=%llu kp=%llu)", (int)r,
                                      (unsigned long long)rows, (unsigned long long)kp);
    return B200KNN_OK;
}

template <typename T>
struct DevBuf {   // grow-only device buffer
    T *p = nullptr;
    size_t cap = 0;
    int ensure(size_t n) {
        if (n <= cap) return B200KNN_OK;
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
        CU_TRY(cudaMalloc(reinterpret_cast<void **>(&p), n * sizeof(T)));
        cap = n;
        return B200KNN_OK;
    }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
};

// Pinned staging ring for pageable host sources, one per device for the whole process: page-locking 96 MB costs
// 30-100 ms, too much to repeat for every handle (the precision/recall metric builds two indexes per evaluation).
struct PinnedRing {
    static constexpr int RING = 3;
    static constexpr size_t BYTES = 32u << 20;
    unsigned char *buf[RING] = {nullptr, nullptr, nullptr};
    cudaEvent_t done[RING] = {nullptr, nullptr, nullptr};
    bool used[RING] = {false, false, false};
    int next = 0;
    std::mutex mu;
};
PinnedRing g_rings[64];

// converted query rows handed to a tensor pass: BF16 rows (+ lo rows for the split tier, or TF32 rows), norms, errors
struct QuerySide {
    const __nv_bfloat16 *bf; const float *norm; const float *err;
    const __nv_bfloat16 *lo = nullptr; const float *tf = nullptr; const float *lonorm = nullptr;
};

}  // namespace
