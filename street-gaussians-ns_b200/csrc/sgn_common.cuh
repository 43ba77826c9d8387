// Shared host/device helpers for libsgn_raster.so (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/sgn_raster.h"

// ---- error plumbing (thread-local message, negative status; nothing throws across the ABI) ----
void sgn_set_error(const char* fmt, ...);

#define SGN_CHECK_CUDA(expr)                                                                     \
    do {                                                                                         \
        cudaError_t _e = (expr);                                                                 \
        if (_e != cudaSuccess) {                                                                 \
            sgn_set_error("%s failed at %s:%d: %s", #expr, __FILE__, __LINE__, cudaGetErrorString(_e)); \
            return SGN_ERR_CUDA;                                                                 \
        }                                                                                        \
    } while (0)

// NVTX range per C-ABI call (nsys / ncu --nvtx timelines name the stages); a no-op function-pointer call when no tool is attached
#include <nvtx3/nvToolsExt.h>
struct SgnRange {
    explicit SgnRange(const char* name) { nvtxRangePushA(name); }
    ~SgnRange() { nvtxRangePop(); }
};
#define SGN_RANGE(name) SgnRange sgn_nvtx_range_(name)

#define SGN_REQUIRE(cond, ...)                 \
    do {                                       \
        if (!(cond)) {                         \
            sgn_set_error(__VA_ARGS__);        \
            return SGN_ERR_INVALID;            \
        }                                      \
    } while (0)

// every kernel launch this library makes is counted (bench.py reports it as gpu_launches); a CUB
// device-wide call counts as one
void sgn_count_launch(int n);

#define SGN_CHECK_LAUNCH(name)                                                                   \
    do {                                                                                         \
        sgn_count_launch(1);                                                                     \
        cudaError_t _e = cudaGetLastError();                                                     \
        if (_e != cudaSuccess) {                                                                 \
            sgn_set_error("launch of %s failed: %s", name, cudaGetErrorString(_e));              \
            return SGN_ERR_CUDA;                                                                 \
        }                                                                                        \
    } while (0)

static __host__ __device__ inline bool sgn_aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// aux word of the projected record
#define SGN_AUX_CLAMP_MASK 0x7
#define SGN_AUX_OBJECT 0x8
#define SGN_AUX_VISIBLE 0x10

#define SGN_TILE 16  // the fused blend kernels are specialised for block_width 16
