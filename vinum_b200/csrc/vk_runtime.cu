// Runtime plumbing of the C ABI: errors, device/stream/event helpers, the
// stream-ordered allocator, and the synthetic-table generator.
#include "vk_common.cuh"
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>

namespace vk {

static thread_local std::string t_last_error;
std::atomic<uint64_t> g_launches{0};

void set_error(const std::string& msg) { t_last_error = msg; }
int fail(int code, const std::string& msg) {
    t_last_error = msg;
    return code;
}
int cuda_fail(cudaError_t e, const char* what) {
    t_last_error = std::string(what) + ": " + cudaGetErrorName(e) + " (" + cudaGetErrorString(e) + ")";
    // clear the sticky-less error state so later calls report their own failures
    cudaGetLastError();
    return e == cudaErrorMemoryAllocation ? VK_ERR_OOM : VK_ERR_CUDA;
}

// ---------------------------------------------------------------- options ----
struct OptEntry { const char* name; int64_t dflt; std::atomic<int64_t> value; std::atomic<int> state; };  // state 0: unread
static OptEntry g_opts[OPT_COUNT] = {
    {"FILTER_STAGE", 1, {0}, {0}},  {"FILTER_PF", 2, {0}, {0}},        
    {"CMP_FAST", 2, {0}, {0}},      {"ARITH_FAST", 4, {0}, {0}},       {"ONEGROUP_FAST", 2, {0}, {0}},
    {"SORT_FUSE_LAST", 1, {0}, {0}}, {"SORT_PREP", 4, {0}, {0}},
    {"AGG_LOG2S", 12, {0}, {0}},    {"AGG_PF", -1, {0}, {0}},          {"AGG_WARPS", 0, {0}, {0}},
    {"AGG_DIRECT", 1, {0}, {0}},    {"AGG_DICT", 1, {0}, {0}},         {"AGG_ENTRY", 1, {0}, {0}},         {"AGG_HOT", 1, {0}, {0}},          {"AGG_NOFAST", 0, {0}, {0}},
    {"AGG_WIDE", 1, {0}, {0}},      {"AGG_PARTITION", 1, {0}, {0}},
    {"AGG_LEARN_LOG2", 20, {0}, {0}}, {"LIST_LOG2", 30, {0}, {0}},     {"DEBUG", 0, {0}, {0}},
    {"INGEST_STAGED", 1, {0}, {0}}, {"INGEST_THREADS", 0, {0}, {0}}, {"INGEST_PIECE_KB", 2048, {0}, {0}},
};
int64_t opt(int id) {
    OptEntry& e = g_opts[id];
    if (e.state.load(std::memory_order_acquire) == 0) {
        const std::string env = std::string("VINUM_B200_") + e.name;
        const char* v = getenv(env.c_str());
        int expected = 0;
        const int64_t val = v && *v ? atoll(v) : e.dflt;
        // a concurrent vk_set_option wins: only an unread entry takes the environment's value
        if (e.state.compare_exchange_strong(expected, 1)) e.value.store(val, std::memory_order_release);
    }
    return e.value.load(std::memory_order_acquire);
}
static int opt_find(const char* name) {
    if (!name) return -1;
    if (strncmp(name, "VINUM_B200_", 11) == 0) name += 11;
    for (int i = 0; i < OPT_COUNT; ++i)
        if (strcmp(name, g_opts[i].name) == 0) return i;
    return -1;
}

static int g_sm_count[64] = {0};
static int g_smem_optin[64] = {0};
int sm_count() {
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64) return 148;
    if (g_sm_count[dev] == 0) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
        g_sm_count[dev] = n;
    }
    return g_sm_count[dev];
}
int max_smem_optin() {
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64) return 227 * 1024;
    if (g_smem_optin[dev] == 0) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev) != cudaSuccess || n <= 0)
            n = 227 * 1024;
        g_smem_optin[dev] = n;
    }
    return g_smem_optin[dev];
}

// --------------------------------------------------------------- datagen ----
// value(r, c) = f_c(splitmix64(r*16 + c + seed*GOLDEN)).  Restated in NumPy in
// vinum_b200/datagen.py; both sides use only exactly-rounded IEEE operations so the
// columns are bit-identical (compiled with -fmad=false).
__device__ __forceinline__ uint64_t gen_u(uint64_t seed, int64_t r, int c) {
    return splitmix64((uint64_t) r * 16ULL + (uint64_t) c + seed * 0x9E3779B97F4A7C15ULL);
}

template <int KIND>
__global__ void __launch_bounds__(256) datagen_kernel(uint64_t seed, int64_t row0, int64_t n, void* out) {
    int64_t stride = (int64_t) gridDim.x * blockDim.x;
    for (int64_t i = (int64_t) blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        int64_t r = row0 + i;
        uint64_t u = gen_u(seed, r, KIND);
        if (KIND == VK_GEN_I0) reinterpret_cast<int64_t*>(out)[i] = (int64_t)(u % 1000ULL);
        else if (KIND == VK_GEN_I1) reinterpret_cast<int64_t*>(out)[i] = (int64_t)(u >> 23) - (1LL << 40);
        else if (KIND == VK_GEN_I2) reinterpret_cast<int64_t*>(out)[i] = r;
        else if (KIND == VK_GEN_I3) reinterpret_cast<int64_t*>(out)[i] = (int64_t)(u % 1000000ULL);
        else if (KIND == VK_GEN_K32) reinterpret_cast<int32_t*>(out)[i] = (int32_t)(u % 1000ULL);
        else {
            double x = (double) (u >> 11) * (1.0 / 9007199254740992.0);  // [0,1), exact
            double v;
            if (KIND == VK_GEN_F0) v = x;
            else if (KIND == VK_GEN_F1) v = (x - 0.5) * 2000.0;
            else if (KIND == VK_GEN_F2) {
                double s = (double) (u & 0xffff) + (double) ((u >> 16) & 0xffff) +
                           (double) ((u >> 32) & 0xffff) + (double) ((u >> 48) & 0xffff);
                v = s * (1.0 / 65536.0) - 2.0;
            } else v = x * 1000000.0;
            reinterpret_cast<double*>(out)[i] = v;
        }
    }
}

}  // namespace vk

using namespace vk;

extern "C" {

int vk_abi_version(void) { return VK_ABI_VERSION; }
const char* vk_last_error(void) { return t_last_error.c_str(); }
uint64_t vk_launch_count(void) { return g_launches.load(); }

int vk_set_option(const char* name, int64_t value) {
    const int i = opt_find(name);
    if (i < 0) return fail(VK_ERR_ARG, std::string("vk_set_option: unknown option '") + (name ? name : "(null)") + "'");
    g_opts[i].value.store(value, std::memory_order_release);
    g_opts[i].state.store(2, std::memory_order_release);
    return VK_OK;
}
int vk_get_option(const char* name, int64_t* out_value) {
    const int i = opt_find(name);
    if (i < 0 || !out_value) return fail(VK_ERR_ARG, std::string("vk_get_option: unknown option '") + (name ? name : "(null)") + "'");
    *out_value = opt(i);
    return VK_OK;
}
int vk_reset_options(void) {
    for (int i = 0; i < OPT_COUNT; ++i) g_opts[i].state.store(0, std::memory_order_release);
    return VK_OK;
}

int vk_device_count(int* out_n) {
    VK_REQUIRE(out_n, "vk_device_count: out_n is NULL");
    VK_CUDA(cudaGetDeviceCount(out_n));
    return VK_OK;
}
int vk_set_device(int device) {
    VK_CUDA(cudaSetDevice(device));
    // keep freed blocks cached in the default pool: operators allocate per batch
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
        uint64_t thr = UINT64_MAX;
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
    }
    return VK_OK;
}
int vk_get_device(int* out_device) {
    VK_REQUIRE(out_device, "vk_get_device: out_device is NULL");
    VK_CUDA(cudaGetDevice(out_device));
    return VK_OK;
}
int vk_device_info(int device, int* out_sm_count, int* out_cc_major, int* out_cc_minor,
                   uint64_t* out_total_bytes, uint64_t* out_free_bytes) {
    cudaDeviceProp p;
    VK_CUDA(cudaGetDeviceProperties(&p, device));
    if (out_sm_count) *out_sm_count = p.multiProcessorCount;
    if (out_cc_major) *out_cc_major = p.major;
    if (out_cc_minor) *out_cc_minor = p.minor;
    if (out_total_bytes || out_free_bytes) {
        int cur = 0;
        VK_CUDA(cudaGetDevice(&cur));
        if (cur != device) VK_CUDA(cudaSetDevice(device));
        size_t f = 0, t = 0;
        VK_CUDA(cudaMemGetInfo(&f, &t));
        if (cur != device) VK_CUDA(cudaSetDevice(cur));
        if (out_total_bytes) *out_total_bytes = t;
        if (out_free_bytes) *out_free_bytes = f;
    }
    return VK_OK;
}

int vk_malloc(void** out_ptr, uint64_t bytes, VkStream stream) {
    VK_REQUIRE(out_ptr, "vk_malloc: out_ptr is NULL");
    *out_ptr = nullptr;
    if (bytes == 0) bytes = 16;
    cudaError_t e = cudaMallocAsync(out_ptr, bytes, (cudaStream_t) stream);
    if (e != cudaSuccess) {
        *out_ptr = nullptr;
        return cuda_fail(e, "cudaMallocAsync");
    }
    return VK_OK;
}
int vk_free(void* ptr, VkStream stream) {
    if (!ptr) return VK_OK;
    VK_CUDA(cudaFreeAsync(ptr, (cudaStream_t) stream));
    return VK_OK;
}
int vk_host_alloc(void** out_ptr, uint64_t bytes) {
    VK_REQUIRE(out_ptr, "vk_host_alloc: out_ptr is NULL");
    *out_ptr = nullptr;
    if (bytes == 0) bytes = 16;
    cudaError_t e = cudaHostAlloc(out_ptr, bytes, cudaHostAllocDefault);
    if (e != cudaSuccess) {
        *out_ptr = nullptr;
        return cuda_fail(e, "cudaHostAlloc");
    }
    return VK_OK;
}
int vk_host_free(void* ptr) {
    if (!ptr) return VK_OK;
    VK_CUDA(cudaFreeHost(ptr));
    return VK_OK;
}
int vk_host_register(void* ptr, uint64_t bytes) {
    VK_REQUIRE(ptr && bytes, "vk_host_register: empty range");
    VK_CUDA(cudaHostRegister(ptr, bytes, cudaHostRegisterDefault));
    return VK_OK;
}
int vk_host_unregister(void* ptr) {
    VK_CUDA(cudaHostUnregister(ptr));
    return VK_OK;
}
int vk_memcpy_h2d(void* dst, const void* host_src, uint64_t bytes, VkStream stream) {
    if (bytes == 0) return VK_OK;
    VK_CUDA(cudaMemcpyAsync(dst, host_src, bytes, cudaMemcpyHostToDevice, (cudaStream_t) stream));
    return VK_OK;
}
int vk_memcpy_d2h(void* host_dst, const void* src, uint64_t bytes, VkStream stream) {
    if (bytes == 0) return VK_OK;
    VK_CUDA(cudaMemcpyAsync(host_dst, src, bytes, cudaMemcpyDeviceToHost, (cudaStream_t) stream));
    return VK_OK;
}
int vk_memcpy_d2d(void* dst, const void* src, uint64_t bytes, VkStream stream) {
    if (bytes == 0) return VK_OK;
    VK_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, (cudaStream_t) stream));
    return VK_OK;
}
int vk_memset(void* dst, int byte, uint64_t bytes, VkStream stream) {
    if (bytes == 0) return VK_OK;
    VK_CUDA(cudaMemsetAsync(dst, byte, bytes, (cudaStream_t) stream));
    return VK_OK;
}
int vk_stream_create(VkStream* out_stream) {
    VK_REQUIRE(out_stream, "vk_stream_create: out_stream is NULL");
    cudaStream_t s;
    VK_CUDA(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
    *out_stream = s;
    return VK_OK;
}
int vk_stream_destroy(VkStream stream) {
    if (stream) VK_CUDA(cudaStreamDestroy((cudaStream_t) stream));
    return VK_OK;
}
int vk_stream_sync(VkStream stream) {
    VK_CUDA(cudaStreamSynchronize((cudaStream_t) stream));
    return VK_OK;
}
int vk_device_sync(void) {
    VK_CUDA(cudaDeviceSynchronize());
    return VK_OK;
}
int vk_event_create(VkEvent* out_event) {
    VK_REQUIRE(out_event, "vk_event_create: out_event is NULL");
    cudaEvent_t e;
    VK_CUDA(cudaEventCreate(&e));
    *out_event = e;
    return VK_OK;
}
int vk_event_destroy(VkEvent event) {
    if (event) VK_CUDA(cudaEventDestroy((cudaEvent_t) event));
    return VK_OK;
}
int vk_event_record(VkEvent event, VkStream stream) {
    VK_CUDA(cudaEventRecord((cudaEvent_t) event, (cudaStream_t) stream));
    return VK_OK;
}
int vk_event_sync(VkEvent event) {
    VK_CUDA(cudaEventSynchronize((cudaEvent_t) event));
    return VK_OK;
}
int vk_stream_wait_event(VkStream stream, VkEvent event) {
    VK_CUDA(cudaStreamWaitEvent((cudaStream_t) stream, (cudaEvent_t) event, 0));
    return VK_OK;
}
int vk_event_elapsed_ms(VkEvent start, VkEvent stop, float* out_ms) {
    VK_REQUIRE(out_ms, "vk_event_elapsed_ms: out_ms is NULL");
    VK_CUDA(cudaEventElapsedTime(out_ms, (cudaEvent_t) start, (cudaEvent_t) stop));
    return VK_OK;
}

int vk_datagen(int kind, uint64_t seed, int64_t row0, int64_t nrows, void* out, VkStream stream) {
    VK_REQUIRE(nrows >= 0, "vk_datagen: negative nrows");
    if (nrows == 0) return VK_OK;
    VK_REQUIRE(out, "vk_datagen: out is NULL");
    int blocks = sm_count() * 8;
    int64_t need = (nrows + 255) / 256;
    if (need < blocks) blocks = (int) need;
    cudaStream_t s = (cudaStream_t) stream;
#define VK_GEN_CASE(K) case K: datagen_kernel<K><<<blocks, 256, 0, s>>>(seed, row0, nrows, out); break;
    switch (kind) {
        VK_GEN_CASE(VK_GEN_I0) VK_GEN_CASE(VK_GEN_I1) VK_GEN_CASE(VK_GEN_I2) VK_GEN_CASE(VK_GEN_I3)
        VK_GEN_CASE(VK_GEN_F0) VK_GEN_CASE(VK_GEN_F1) VK_GEN_CASE(VK_GEN_F2) VK_GEN_CASE(VK_GEN_F3)
        VK_GEN_CASE(VK_GEN_K32)
        default: return fail(VK_ERR_ARG, "vk_datagen: unknown column kind");
    }
#undef VK_GEN_CASE
    VK_CHECK_LAUNCH("datagen_kernel");
    return VK_OK;
}

}  // extern "C"
