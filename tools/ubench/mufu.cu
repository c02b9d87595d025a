// MUFU throughput microbenchmark (sm_100a): results per clock per SM for fp32 / packed 16-bit transcendental ops.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mufu mufu.cu && ./mufu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int OP>
__global__ void k(uint32_t* out, int iters) {
    uint32_t a[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = 0x3c003c00u + threadIdx.x * 8 + i;   // ~1.0 in f16x2 / small floats
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (OP == 0) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+r"(a[i]));
            if (OP == 1) asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(a[i]));
            if (OP == 2) asm volatile("ex2.approx.ftz.bf16x2 %0, %0;" : "+r"(a[i]));
            if (OP == 3) asm volatile("tanh.approx.f32 %0, %0;" : "+r"(a[i]));
            if (OP == 4) asm volatile("tanh.approx.f16x2 %0, %0;" : "+r"(a[i]));
            if (OP == 5) asm volatile("tanh.approx.bf16x2 %0, %0;" : "+r"(a[i]));
            if (OP == 6) asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+r"(a[i]));
            if (OP == 7) asm volatile("fma.rn.f32 %0, %0, %0, %0;" : "+r"(a[i]));
            if (OP == 8) asm volatile("fma.rn.bf16x2 %0, %0, %0, %0;" : "+r"(a[i]));
        }
    }
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s ^= a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int OP>
void run(const char* name, int elems_per_op) {
    uint32_t* d; cudaMalloc(&d, 148 * 8 * 1024 * 4);
    const int iters = 4096, blocks = 148 * 2, threads = 1024;
    k<OP><<<blocks, threads>>>(d, 16);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k<OP><<<blocks, threads>>>(d, iters);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    int clk_khz; cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
    const double ops = (double)blocks * threads * iters * 8;
    const double per_clk_sm = ops / (ms * 1e-3) / (clk_khz * 1e3) / 148.0;
    printf("%-24s %8.3f ms  %7.2f lane-ops/clk/SM (at max clock %d MHz)  -> %7.2f results/clk/SM\n", name, ms, per_clk_sm,
           clk_khz / 1000, per_clk_sm * elems_per_op);
    cudaFree(d);
}

int main() {
    run<0>("ex2.approx.ftz.f32", 1);
    run<1>("ex2.approx.f16x2", 2);
    run<2>("ex2.approx.ftz.bf16x2", 2);
    run<3>("tanh.approx.f32", 1);
    run<4>("tanh.approx.f16x2", 2);
    run<5>("tanh.approx.bf16x2", 2);
    run<6>("rcp.approx.ftz.f32", 1);
    run<7>("fma.rn.f32", 1);
    run<8>("fma.rn.bf16x2", 2);
    return 0;
}
