// Yardstick only -- never linked into libvinum_b200.so (SURVEY Appendix B allows CUB as a
// comparison baseline).  Times, on the same box as our kernels:
//   * cub::DeviceRadixSort::SortPairs<uint64 key, uint32 idx> at 1e8 keys (C4's core),
//   * cub::DeviceSelect::Flagged over 4 x 8-byte columns at 1e8 rows, sel 0.5 (C2's core; the
//     byte flags are produced by a separate compare kernel, as a library user would do),
//   * a plain device-to-device copy (the "peak" both are judged against).
//
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o gpurun_out/cub_baseline scripts/ubench/cub_baseline.cu
#include <cstdint>
#include <cstdio>
#include <cub/cub.cuh>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

__device__ __host__ inline uint64_t splitmix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ULL;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ULL;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBULL;
    return x ^ (x >> 31);
}
__global__ void gen_kernel(uint64_t* k, uint32_t* idx, double* f, int64_t n) {
    for (int64_t i = blockIdx.x * (int64_t) blockDim.x + threadIdx.x; i < n; i += (int64_t) gridDim.x * blockDim.x) {
        const uint64_t u = splitmix64(i * 8 + 7);
        if (k) k[i] = u;
        if (idx) idx[i] = (uint32_t) i;
        if (f) f[i] = (double) (u >> 11) * (1.0 / 9007199254740992.0);
    }
}
__global__ void flag_kernel(const double* f, uint8_t* flag, int64_t n, double c) {
    for (int64_t i = blockIdx.x * (int64_t) blockDim.x + threadIdx.x; i < n; i += (int64_t) gridDim.x * blockDim.x) flag[i] = f[i] > c;
}

template <class F>
float best_ms(F fn, int reps = 5) {
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    fn(); fn();
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int i = 0; i < reps; ++i) {
        cudaEventRecord(a);
        fn();
        cudaEventRecord(b);
        cudaEventSynchronize(b);
        float ms; cudaEventElapsedTime(&ms, a, b);
        best = ms < best ? ms : best;
    }
    return best;
}

int main() {
    const int64_t n = 100000000;
    // ---- radix sort pairs ----
    uint64_t *k0, *k1; uint32_t *v0, *v1;
    CK(cudaMalloc(&k0, n * 8)); CK(cudaMalloc(&k1, n * 8)); CK(cudaMalloc(&v0, n * 4)); CK(cudaMalloc(&v1, n * 4));
    gen_kernel<<<148 * 8, 256>>>(k0, v0, nullptr, n);
    size_t tb = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, tb, k0, k1, v0, v1, n);
    void* tmp; CK(cudaMalloc(&tmp, tb));
    float ms = best_ms([&] { cub::DeviceRadixSort::SortPairs(tmp, tb, k0, k1, v0, v1, n); });
    printf("{\"cub_sort_pairs_u64_u32_1e8_ms\": %.4f, \"grows_per_s\": %.3f}\n", ms, n / ms / 1e6);
    size_t tbk = 0;
    cub::DeviceRadixSort::SortKeys(nullptr, tbk, k0, k1, n);
    void* tmpk; CK(cudaMalloc(&tmpk, tbk));
    ms = best_ms([&] { cub::DeviceRadixSort::SortKeys(tmpk, tbk, k0, k1, n); });
    printf("{\"cub_sort_keys_u64_1e8_ms\": %.4f, \"grows_per_s\": %.3f}\n", ms, n / ms / 1e6);
    CK(cudaFree(tmp)); CK(cudaFree(tmpk)); CK(cudaFree(v0)); CK(cudaFree(v1));
    // ---- copy peak ----
    ms = best_ms([&] { cudaMemcpyAsync(k1, k0, n * 8, cudaMemcpyDeviceToDevice); });
    printf("{\"d2d_copy_0.8GB_ms\": %.4f, \"GBps\": %.1f}\n", ms, 2.0 * n * 8 / ms / 1e6);
    CK(cudaFree(k0)); CK(cudaFree(k1));
    // ---- select flagged, 4 columns ----
    double* cols[4]; double* outs[4]; uint8_t* flag; int64_t* nsel;
    for (int c = 0; c < 4; ++c) { CK(cudaMalloc(&cols[c], n * 8)); CK(cudaMalloc(&outs[c], n * 8)); gen_kernel<<<148 * 8, 256>>>(nullptr, nullptr, cols[c], n); }
    CK(cudaMalloc(&flag, n)); CK(cudaMalloc(&nsel, 8));
    size_t ts = 0;
    cub::DeviceSelect::Flagged(nullptr, ts, cols[0], flag, outs[0], nsel, n);
    void* tmps; CK(cudaMalloc(&tmps, ts));
    ms = best_ms([&] {
        flag_kernel<<<148 * 8, 256>>>(cols[2], flag, n, 0.5);
        for (int c = 0; c < 4; ++c) cub::DeviceSelect::Flagged(tmps, ts, cols[c], flag, outs[c], nsel, n);
    });
    printf("{\"cub_select_flagged_4cols_1e8_ms\": %.4f, \"GBps_algorithmic_48B\": %.1f}\n", ms, 48.0 * n / ms / 1e6);
    ms = best_ms([&] { cub::DeviceSelect::Flagged(tmps, ts, cols[0], flag, outs[0], nsel, n); });
    printf("{\"cub_select_flagged_1col_1e8_ms\": %.4f}\n", ms);
    return 0;
}
