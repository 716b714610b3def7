// ubench_smem.cu — what a B200 SM sustains for the shared-memory operations the pileup kernels lean on:
// red.shared.or (no return), st.shared and ld.shared, 32 lanes on 32 different banks (rows of 33 words) and on
// random words of the warp's slice.  Prints lane-operations per clock per SM.  nvcc -arch=sm_100a -O3.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ void reds_or(uint32_t a, uint32_t v) { asm volatile("red.shared.or.b32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ void sts(uint32_t a, uint32_t v) { asm volatile("st.shared.u32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ uint32_t lds(uint32_t a) { uint32_t v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a) : "memory"); return v; }

constexpr int WORDS = 1056;         // a warp's rows
template <int MODE, int PATTERN>
__global__ void k(int iters, unsigned long long* out, uint32_t* sink) {
    extern __shared__ uint32_t sm[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t base = (uint32_t)__cvta_generic_to_shared(sm) + 4u * warp * WORDS;
    for (int i = lane; i < WORDS; i += 32) sts(base + 4 * i, 0u);
    __syncthreads();
    uint32_t acc = 0, h = lane * 2654435761u + warp;
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            uint32_t idx;
            if (PATTERN == 0) idx = (uint32_t)(lane * 33 + ((it * 8 + u) & 31));          // one bank per lane
            else { h = h * 1664525u + 1013904223u; idx = (h >> 8) % WORDS; }              // random words
            const uint32_t a = base + 4u * idx;
            if (MODE == 0) reds_or(a, 1u << u);
            else if (MODE == 1) sts(a, h);
            else acc += lds(a);
        }
    }
    const long long t1 = clock64();
    __syncthreads();
    if (threadIdx.x == 0) out[blockIdx.x] = (unsigned long long)(t1 - t0);
    if (acc == 0x12345) sink[0] = acc;
}

template <int MODE, int PATTERN>
void run(const char* name, int warps) {
    unsigned long long* d; uint32_t* sink;
    cudaMalloc(&d, 148 * 8); cudaMalloc(&sink, 4);
    const int iters = 4000;
    const size_t smem = 4 * WORDS * warps;
    cudaFuncSetAttribute(k<MODE, PATTERN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    k<MODE, PATTERN><<<148, warps * 32, smem>>>(10, d, sink);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k<MODE, PATTERN><<<148, warps * 32, smem>>>(iters, d, sink);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    unsigned long long h[148]; cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    double cyc = 0; for (int i = 0; i < 148; ++i) cyc += (double)h[i]; cyc /= 148;
    const double laneops = (double)iters * 8 * 32 * warps;
    printf("%-34s warps %2d: %.2f lane-ops/clk/SM (%.0f cycles, %.3f ms) err=%d\n", name, warps, laneops / cyc, cyc, ms, (int)cudaGetLastError());
    cudaFree(d); cudaFree(sink);
}

int main() {
    for (int warps : {4, 8, 16}) {
        run<0, 0>("red.shared.or  one bank per lane", warps);
        run<0, 1>("red.shared.or  random words", warps);
        run<1, 0>("st.shared      one bank per lane", warps);
        run<1, 1>("st.shared      random words", warps);
        run<2, 0>("ld.shared      one bank per lane", warps);
        run<2, 1>("ld.shared      random words", warps);
    }
    return 0;
}
