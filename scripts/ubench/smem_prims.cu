// Shared-memory primitive costs on sm_100a, as seen by a group-by kernel: random-address
// LDS/STS of 4/8/16 bytes, shared atomics, MATCH.ANY, and the tag-arbitration round.
// Prints cycles per warp-instruction per SM (SM-wide reciprocal throughput).
//
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o gpurun_out/smem_prims scripts/ubench/smem_prims.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

constexpr int G = 1024;  // distinct slots addressed
constexpr int ITERS = 4096;
constexpr int U = 8;    // independent accesses in flight per warp

__device__ __forceinline__ uint32_t next_idx(uint32_t& s) {
    s = s * 1664525u + 1013904223u;
    return (s >> 10) & (G - 1);
}

enum Prim { P_NOP, P_LDS32, P_LDS64, P_LDS128, P_STS32, P_STS64, P_STS128, P_ATOMS32, P_ATOMS32_RET, P_CAS64,
            P_ATOMADD64, P_MATCH, P_TAG8, P_RMW128, P_LDS16, P_ATOMF64, P_COUNT };
const char* NAMES[] = {"nop(idx gen only)", "LDS.32 random", "LDS.64 random", "LDS.128 random", "STS.32 random",
                       "STS.64 random", "STS.128 random", "ATOMS.ADD.32 noret", "ATOMS.ADD.32 ret", "ATOMS.CAS.64",
                       "atomicAdd u64 smem", "MATCH.ANY", "tag8 STS+syncwarp+LDS", "LDS.128+STS.128 RMW", "LDS.U16 random",
                       "atomicAdd f64 smem"};

template <int P>
__global__ void __launch_bounds__(512) prim_kernel(uint64_t* out, int active_mod) {
    extern __shared__ __align__(16) uint8_t smem[];
    uint32_t* s32 = reinterpret_cast<uint32_t*>(smem);
    uint64_t* s64 = reinterpret_cast<uint64_t*>(smem);
    uint4* s128 = reinterpret_cast<uint4*>(smem);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // each warp owns a private 16 KB region (G x 16 B)
    const int woff = warp * G;
    for (int i = threadIdx.x; i < (int) (blockDim.x / 32) * G * 4; i += blockDim.x) s32[i] = 0;
    __syncthreads();
    uint32_t seed = (blockIdx.x * blockDim.x + threadIdx.x) * 2654435761u + 12345u;
    uint64_t acc = 0;
    const bool on = (lane % active_mod) == 0;  // active_mod = 2 -> 16 active lanes (selectivity 0.5)
    for (int it = 0; it < ITERS / U; ++it) {
      uint32_t rnd[U];
      uint64_t part[U];
#pragma unroll
      for (int u = 0; u < U; ++u) { rnd[u] = next_idx(seed); part[u] = 0; }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const uint32_t idx = rnd[u];
        uint64_t& acc = part[u];
        if constexpr (P == P_NOP) acc += idx;
        if constexpr (P == P_LDS32) { if (on) acc += reinterpret_cast<volatile uint32_t*>(s32)[woff * 4 + idx]; }
        if constexpr (P == P_LDS16) { if (on) acc += reinterpret_cast<volatile uint16_t*>(s32)[woff * 8 + idx]; }
        if constexpr (P == P_LDS64) { if (on) acc += reinterpret_cast<volatile uint64_t*>(s64)[woff * 2 + idx]; }
        if constexpr (P == P_LDS128) {
            if (on) {
                uint32_t a, b, c, d;
                uint32_t addr = (uint32_t) __cvta_generic_to_shared(s128 + woff + idx);
                asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(a), "=r"(b), "=r"(c), "=r"(d) : "r"(addr) : "memory");
                acc += a + b + c + d;
            }
        }
        if constexpr (P == P_STS32) { if (on) reinterpret_cast<volatile uint32_t*>(s32)[woff * 4 + idx] = it; }
        if constexpr (P == P_STS64) { if (on) reinterpret_cast<volatile uint64_t*>(s64)[woff * 2 + idx] = it; }
        if constexpr (P == P_STS128) {
            if (on) {
                uint32_t addr = (uint32_t) __cvta_generic_to_shared(s128 + woff + idx);
                asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(it), "r"(it), "r"(it), "r"(it) : "memory");
            }
        }
        if constexpr (P == P_ATOMS32) { if (on) atomicAdd(&s32[woff * 4 + idx], 1u); }
        if constexpr (P == P_ATOMS32_RET) { if (on) acc += atomicAdd(&s32[woff * 4 + idx], 1u); }
        if constexpr (P == P_CAS64) {
            if (on) acc += atomicCAS(reinterpret_cast<unsigned long long*>(s64 + woff * 2 + idx), (unsigned long long) it, (unsigned long long) it + 1);
        }
        if constexpr (P == P_ATOMADD64) { if (on) atomicAdd(reinterpret_cast<unsigned long long*>(s64 + woff * 2 + idx), 1ULL); }
        if constexpr (P == P_ATOMF64) { if (on) atomicAdd(reinterpret_cast<double*>(s64 + woff * 2 + idx), 1.0); }
        if constexpr (P == P_MATCH) { acc += __match_any_sync(0xffffffffu, on ? idx : (0x10000u | lane)); }
      }
      if constexpr (P == P_TAG8) {  // batch form: U tag stores, sync, U tag loads, sync
          volatile uint8_t* tags = reinterpret_cast<volatile uint8_t*>(smem) + woff * 16;
#pragma unroll
          for (int u = 0; u < U; ++u) if (on) tags[rnd[u]] = (uint8_t) lane;
          __syncwarp();
#pragma unroll
          for (int u = 0; u < U; ++u) if (on) part[u] += tags[rnd[u]];
          __syncwarp();
      }
      if constexpr (P == P_RMW128) {  // batch form: U loads, U adds, U stores, sync
          uint32_t a[U], b[U], c[U], d[U];
#pragma unroll
          for (int u = 0; u < U; ++u) if (on) {
              uint32_t addr = (uint32_t) __cvta_generic_to_shared(s128 + woff + rnd[u]);
              asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(a[u]), "=r"(b[u]), "=r"(c[u]), "=r"(d[u]) : "r"(addr));
          }
#pragma unroll
          for (int u = 0; u < U; ++u) if (on) {
              uint32_t addr = (uint32_t) __cvta_generic_to_shared(s128 + woff + rnd[u]);
              double sm = __hiloint2double(b[u], a[u]) + 1.5;
              asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(__double2loint(sm)), "r"(__double2hiint(sm)), "r"(c[u] + 1), "r"(d[u]) : "memory");
          }
          __syncwarp();
      }
#pragma unroll
      for (int u = 0; u < U; ++u) acc += part[u];
    }
    if (acc == 0x123456789abcULL) out[0] = acc;
}

template <int P>
void run(int warps, int active_mod, uint64_t* d_out, int sms) {
    const int threads = warps * 32;
    const size_t smem = (size_t) warps * G * 16;
    cudaFuncSetAttribute(prim_kernel<P>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    prim_kernel<P><<<sms, threads, smem>>>(d_out, active_mod);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    prim_kernel<P><<<sms, threads, smem>>>(d_out, active_mod);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    cudaError_t err = cudaGetLastError();
    int clk = 0;
    cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    // cycles per warp-instruction per SM, at the nominal max clock
    const double cyc = (double) ms * 1e-3 * clk * 1e3 / ((double) ITERS * warps);
    printf("%-26s warps=%2d active=%2d  %8.3f ms  %7.2f cyc/warp-instr/SM %s\n", NAMES[P], warps, 32 / active_mod, ms, cyc,
           err == cudaSuccess ? "" : cudaGetErrorString(err));
}

template <int P>
void run_all(uint64_t* d_out, int sms) {
    for (int warps : {8, 16})
        for (int am : {1, 2}) run<P>(warps, am, d_out, sms);
}

int main() {
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    uint64_t* d_out;
    cudaMalloc(&d_out, 64);
    run_all<P_NOP>(d_out, sms);
    run_all<P_LDS16>(d_out, sms);
    run_all<P_LDS32>(d_out, sms);
    run_all<P_LDS64>(d_out, sms);
    run_all<P_LDS128>(d_out, sms);
    run_all<P_STS32>(d_out, sms);
    run_all<P_STS64>(d_out, sms);
    run_all<P_STS128>(d_out, sms);
    run_all<P_RMW128>(d_out, sms);
    run_all<P_TAG8>(d_out, sms);
    run_all<P_MATCH>(d_out, sms);
    run_all<P_ATOMS32>(d_out, sms);
    run_all<P_ATOMS32_RET>(d_out, sms);
    run_all<P_CAS64>(d_out, sms);
    run_all<P_ATOMADD64>(d_out, sms);
    run_all<P_ATOMF64>(d_out, sms);
    return 0;
}
