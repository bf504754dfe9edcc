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

// Per-DEVICE one-time state (runtime.cu).  Several engines on different GPUs may live in one process and be driven from
// different host threads: nothing below is cached process-wide.
//   ensure_kernel_smem : cudaFuncSetAttribute(MaxDynamicSharedMemorySize [+ carveout]) once per (device, kernel)
//   device_sm_count    : multiprocessor count of the current device
//   device_first_use   : true exactly once per (current device, key): guard for lazily initialised device tables
int ensure_kernel_smem(const void* func, int bytes, bool max_carveout = false);
int device_sm_count(int* sms);
int device_first_use(const void* key, bool* first);

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

// RISE keep mask (src/spectrogram_explainability.py:768: an i.i.d. Bernoulli(p) bit per spectrogram cell and mask; the
// reference draws it from the UNSEEDED numpy global RNG, so the bit generator here is the builder's: a counter-based hash
// of (seed, mask index, cell) that the iSTFT load stage, the map reduction and the CPU oracle all evaluate identically).
__host__ __device__ inline uint32_t hash_lowbias32(uint32_t x) {
    x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16;
    return x;
}
__host__ __device__ inline uint32_t rise_mask_key(uint32_t seed, uint32_t mask_index) {
    return hash_lowbias32(seed * 0x9E3779B9U + mask_index * 0x85EBCA6BU + 0x165667B1U);
}
// cell = frame * 1025 + bin; kept iff the 32-bit uniform is below floor(p * 2^32)
__host__ __device__ inline bool rise_keep(uint32_t key, uint32_t cell, uint32_t threshold) {
    return hash_lowbias32(key ^ (cell * 0xC2B2AE35U)) < threshold;
}
inline uint32_t rise_threshold(double keep_probability) {
    const double v = keep_probability * 4294967296.0;
    return v >= 4294967295.0 ? 4294967295U : (v <= 0.0 ? 0U : static_cast<uint32_t>(v));
}

}  // namespace b200x
