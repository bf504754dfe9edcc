// Library-wide runtime pieces of the C-ABI: last-error string, version, tensor-map creation.
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <map>
#include <mutex>
#include <set>
#include <unordered_map>
#include <utility>

#include "common.h"

namespace b200x {

static thread_local char g_err[1024] = "";

// ---- per-device one-time state ----------------------------------------------------------------------------------------
static std::mutex g_dev_mutex;
static std::set<std::pair<int, const void*>> g_dev_done;       // (device, key) pairs already initialised / configured
static std::map<int, int> g_dev_sms;

int device_first_use(const void* key, bool* first) {
    int dev = 0;
    B200X_CUDA_TRY(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lock(g_dev_mutex);
    *first = g_dev_done.insert(std::make_pair(dev, key)).second;
    return B200X_OK;
}

int ensure_kernel_smem(const void* func, int bytes, bool max_carveout) {
    int dev = 0;
    B200X_CUDA_TRY(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lock(g_dev_mutex);             // held across the attribute calls: a second thread must not launch early
    if (g_dev_done.count(std::make_pair(dev, func))) return B200X_OK;
    B200X_CUDA_TRY(cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
    if (max_carveout) B200X_CUDA_TRY(cudaFuncSetAttribute(func, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    g_dev_done.insert(std::make_pair(dev, func));
    return B200X_OK;
}

int device_sm_count(int* sms) {
    int dev = 0;
    B200X_CUDA_TRY(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lock(g_dev_mutex);
    auto it = g_dev_sms.find(dev);
    if (it == g_dev_sms.end()) {
        int n = 0;
        B200X_CUDA_TRY(cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev));
        it = g_dev_sms.emplace(dev, n).first;
    }
    *sms = it->second;
    return B200X_OK;
}

int set_error(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

// cuTensorMapEncodeTiled is resolved through the runtime's driver entry point so that the shared library has
// no link-time dependency on libcuda.so (it must load, for the symbol check, on a box without a driver).
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static PFN_encodeTiled g_encode = nullptr;
static std::once_flag g_encode_once;

// Encoded tensor maps are cached: the engine issues the same GEMM / attention launches (same buffers, same shapes)
// thousands of times per track, and cuTensorMapEncodeTiled costs microseconds of host time per call.
struct TmapKey {
    uint64_t v[11];
    bool operator==(const TmapKey& o) const { return std::memcmp(v, o.v, sizeof(v)) == 0; }
};
struct TmapKeyHash {
    size_t operator()(const TmapKey& k) const {
        uint64_t h = 1469598103934665603ull;
        for (uint64_t x : k.v) { h ^= x; h *= 1099511628211ull; }
        return static_cast<size_t>(h);
    }
};
static std::unordered_map<TmapKey, CUtensorMap, TmapKeyHash> g_tmap_cache;
static std::mutex g_tmap_mutex;

int make_tmap_bf16(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                   const uint32_t* box) {
    return make_tmap(out, base, 2, rank, dims, strides_bytes, box, 1);
}

int make_tmap(CUtensorMap* out, const void* base, int elem_bytes, int rank, const uint64_t* dims,
              const uint64_t* strides_bytes, const uint32_t* box, int swizzle128) {
    std::call_once(g_encode_once, [] {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            g_encode = reinterpret_cast<PFN_encodeTiled>(fn);
    });
    if (g_encode == nullptr) return set_error(B200X_ERR_CUDA, "cuTensorMapEncodeTiled entry point unavailable");
    if ((reinterpret_cast<uintptr_t>(base) & 15) != 0) return set_error(B200X_ERR_INVALID, "tensor map base not 16-byte aligned");
    TmapKey key{};
    key.v[0] = reinterpret_cast<uintptr_t>(base);
    key.v[1] = static_cast<uint64_t>(elem_bytes) | (static_cast<uint64_t>(rank) << 8) | (static_cast<uint64_t>(swizzle128) << 16);
    for (int i = 0; i < rank; ++i) {
        key.v[2 + i] = dims[i];
        key.v[5 + i] = box[i];
        if (i > 0) key.v[8 + i - 1] = strides_bytes[i - 1];
    }
    {
        std::lock_guard<std::mutex> lock(g_tmap_mutex);
        auto it = g_tmap_cache.find(key);
        if (it != g_tmap_cache.end()) { *out = it->second; return B200X_OK; }
    }
    cuuint64_t gdim[3];
    cuuint64_t gstr[2];
    cuuint32_t bdim[3], estr[3];
    for (int i = 0; i < rank; ++i) {
        gdim[i] = dims[i];
        bdim[i] = box[i];
        estr[i] = 1;
        if (i > 0) {
            if (strides_bytes[i - 1] % 16 != 0) return set_error(B200X_ERR_INVALID, "tensor map stride %d not 16-byte aligned", i);
            gstr[i - 1] = strides_bytes[i - 1];
        }
    }
    CUresult r = g_encode(out, elem_bytes == 4 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16,
                          static_cast<cuuint32_t>(rank), const_cast<void*>(base), gdim, gstr, bdim, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE,
                          CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return set_error(B200X_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d", static_cast<int>(r));
    {
        std::lock_guard<std::mutex> lock(g_tmap_mutex);
        if (g_tmap_cache.size() > 8192) g_tmap_cache.clear();
        g_tmap_cache.emplace(key, *out);
    }
    return B200X_OK;
}

}  // namespace b200x

extern "C" const char* b200x_last_error(void) { return b200x::g_err; }
extern "C" int b200x_version(void) { return 100; }
extern "C" int b200x_set_device(int device) {
    B200X_CUDA_TRY(cudaSetDevice(device));
    return B200X_OK;
}
extern "C" int b200x_device_count(int* count) {
    B200X_REQUIRE(count != nullptr, "device_count: NULL output");
    B200X_CUDA_TRY(cudaGetDeviceCount(count));
    return B200X_OK;
}
