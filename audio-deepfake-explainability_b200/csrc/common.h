// Host-side helpers shared by the .cu translation units: error reporting for the C-ABI, CUDA checks,
// and TMA tensor-map creation through the driver entry point (no link-time dependency on libcuda, so
// the library loads on a machine without a GPU driver for the symbol-export check).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <string>

#include "../../include/b200xai.h"

namespace b200x {

int set_error(int code, const char* fmt, ...);

#define B200X_CUDA_TRY(expr)                                                                         \
    do {                                                                                             \
        cudaError_t _e = (expr);                                                                     \
        if (_e != cudaSuccess)                                                                       \
            return ::b200x::set_error(B200X_ERR_CUDA, "%s failed at %s:%d: %s", #expr, __FILE__, __LINE__, \
                                      cudaGetErrorString(_e));                                       \
    } while (0)

#define B200X_TRY(expr)                  \
    do {                                 \
        int _s = (expr);                 \
        if (_s != B200X_OK) return _s;   \
    } while (0)

#define B200X_REQUIRE(cond, ...)                                              \
    do {                                                                      \
        if (!(cond)) return ::b200x::set_error(B200X_ERR_INVALID, __VA_ARGS__); \
    } while (0)

// bf16 tensor map, up to 3 dims (dim0 innermost), 128-byte swizzle, OOB reads filled with zero.
int make_tmap_bf16(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                   const uint32_t* box);
// general form: elem_bytes 2 (bf16) or 4 (fp32); swizzle128 = 0 -> dense (un-swizzled) shared-memory box
int make_tmap(CUtensorMap* out, const void* base, int elem_bytes, int rank, const uint64_t* dims,
              const uint64_t* strides_bytes, const uint32_t* box, int swizzle128);

inline int ceil_div(int a, int b) { return (a + b - 1) / b; }

}  // namespace b200x
