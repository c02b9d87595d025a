// tcgen05.ld (LDTM) throughput microbenchmark (sm_100a): bytes per clock per SM for 1, 2, 4, 8, 16 reading warps, alone and
// interleaved with MUFU.EX2 work (the mix of the attention softmax warps: 1 ex2 + 1 fadd + 0.5 cvt per loaded element).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ldtm ldtm.cu && ./ldtm
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int MODE>   // 0: loads only, 1: loads + exp of every element, 2: exp only (same count, no loads)
__global__ void __launch_bounds__(512, 1) k(uint32_t* out, int iters, int nwarps, long long* clk) {
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = slot;
    uint32_t acc = 0;
    float facc = 0.f;
    const long long t0 = clock64();
    if (warp < nwarps) {
        const uint32_t addr = tmem + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)((warp >> 2) * 64);
        uint32_t v[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = threadIdx.x * 16 + i;
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int p = 0; p < 4; ++p) {
                if (MODE != 2) {
                    asm volatile(
                        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
                        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                        : "r"(addr + p * 16)
                        : "memory");
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                }
                if (MODE == 0) {
#pragma unroll
                    for (int i = 0; i < 16; ++i) acc ^= v[i];
                } else {
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        float x = __uint_as_float(v[i]) - 3.0f, y;
                        asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
                        facc += y;
                        if (MODE == 2) v[i] = __float_as_uint(y * 0.001f);
                    }
                }
            }
        }
    }
    const long long t1 = clock64();
    if (threadIdx.x == 0 && blockIdx.x == 0) *clk = t1 - t0;
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc ^ __float_as_uint(facc);
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
}

template <int MODE>
void run(const char* name, int nwarps) {
    uint32_t* d; cudaMalloc(&d, 148 * 512 * 4);
    long long* clk; cudaMalloc(&clk, 8);
    const int iters = 2000;
    k<MODE><<<148, 512>>>(d, 10, nwarps, clk);
    k<MODE><<<148, 512>>>(d, iters, nwarps, clk);
    cudaError_t e = cudaDeviceSynchronize();
    long long h = 0; cudaMemcpy(&h, clk, 8, cudaMemcpyDeviceToHost);
    const double bytes = (double)nwarps * iters * 4 * 16 * 32 * 4;       // per SM
    const double elems = (double)nwarps * iters * 4 * 16 * 32;
    printf("%-28s warps %2d  %9lld clk  %7.1f B/clk/SM  %6.2f elements/clk/SM  (%s)\n", name, nwarps, h, MODE == 2 ? 0.0 : bytes / h,
           elems / h, cudaGetErrorString(e));
    cudaFree(d); cudaFree(clk);
}

int main() {
    for (int w : {1, 2, 4, 8, 16}) run<0>("LDTM.x16 only", w);
    for (int w : {4, 8, 16}) run<2>("ex2 only", w);
    for (int w : {4, 8, 16}) run<1>("LDTM.x16 + ex2 per element", w);
    return 0;
}
