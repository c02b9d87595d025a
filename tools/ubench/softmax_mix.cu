// Which part of the attention softmax loop costs MUFU throughput?  16 warps per SM (the kernel's 2 CTAs x 8 softmax warps)
// run the per-unit sequence of attention_tc.cu step by step: LDTM.x16 x4 + 64 x (FADD, EX2) [+ F2FP pack] [+ STS.128
// swizzled] [+ fence.proxy.async + mbarrier arrive per unit] [+ 4 sampled FMNMX per piece].
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o softmax_mix softmax_mix.cu && ./softmax_mix
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
    uint32_t r; asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo)); return r;
}

template <int LEVEL>
__global__ void __launch_bounds__(512, 1) k(uint32_t* out, int iters, long long* clk) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint32_t slot;
    __shared__ uint64_t bar[16];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (threadIdx.x < 16) asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar[threadIdx.x])), "r"((1 << 20) - 1));
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = slot;
    const uint32_t addr = tmem + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)((warp >> 2) * 64);
    const int r = (warp & 3) * 32 + lane, sw = r & 7;
    uint8_t* p_row = smem + (warp >> 2) * 16384 + r * 128;
    const float mrow = 3.0f;
    float smp = 0.f;
    uint32_t vv[2][16];
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                     : "=r"(vv[0][0]), "=r"(vv[0][1]), "=r"(vv[0][2]), "=r"(vv[0][3]), "=r"(vv[0][4]), "=r"(vv[0][5]), "=r"(vv[0][6]), "=r"(vv[0][7]),
                       "=r"(vv[0][8]), "=r"(vv[0][9]), "=r"(vv[0][10]), "=r"(vv[0][11]), "=r"(vv[0][12]), "=r"(vv[0][13]), "=r"(vv[0][14]), "=r"(vv[0][15])
                     : "r"(addr) : "memory");
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int piece = 0; piece < 4; ++piece) {
            uint32_t (&cur)[16] = vv[piece & 1];
            uint32_t (&nxt)[16] = vv[(piece + 1) & 1];
            if (piece + 1 < 4)
                asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                             : "=r"(nxt[0]), "=r"(nxt[1]), "=r"(nxt[2]), "=r"(nxt[3]), "=r"(nxt[4]), "=r"(nxt[5]), "=r"(nxt[6]), "=r"(nxt[7]),
                               "=r"(nxt[8]), "=r"(nxt[9]), "=r"(nxt[10]), "=r"(nxt[11]), "=r"(nxt[12]), "=r"(nxt[13]), "=r"(nxt[14]), "=r"(nxt[15])
                             : "r"(addr + (piece + 1) * 16) : "memory");
            if (LEVEL >= 4) {
                smp = fmaxf(fmaxf(smp, __uint_as_float(cur[0])), __uint_as_float(cur[4]));
                smp = fmaxf(fmaxf(smp, __uint_as_float(cur[8])), __uint_as_float(cur[12]));
            }
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                uint32_t w[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int e = q * 8 + i * 2;
                    float y0, y1;
                    asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y0) : "f"(__uint_as_float(cur[e]) - mrow));
                    asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y1) : "f"(__uint_as_float(cur[e + 1]) - mrow));
                    if (LEVEL >= 1) w[i] = pack_bf16x2(y0, y1);
                    else w[i] = __float_as_uint(y0 + y1);
                }
                if (LEVEL >= 2) *reinterpret_cast<uint4*>(p_row + (((piece * 2 + q) ^ sw) << 4)) = make_uint4(w[0], w[1], w[2], w[3]);
                else smp += __uint_as_float(w[0] ^ w[1] ^ w[2] ^ w[3]);
            }
            if (piece + 1 < 4) asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        }
        if (LEVEL >= 3) {
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&bar[warp])) : "memory");
        }
    }
    const long long t1 = clock64();
    if (threadIdx.x == 0 && blockIdx.x == 0) *clk = t1 - t0;
    out[blockIdx.x * blockDim.x + threadIdx.x] = __float_as_uint(smp) ^ smem[threadIdx.x * 16];
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
}

template <int LEVEL>
void run(const char* name) {
    uint32_t* d; cudaMalloc(&d, 148 * 512 * 4);
    long long* clk; cudaMalloc(&clk, 8);
    const int iters = 2000, smem = 4 * 16384;
    cudaFuncSetAttribute(k<LEVEL>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    k<LEVEL><<<148, 512, smem>>>(d, 10, clk);
    k<LEVEL><<<148, 512, smem>>>(d, iters, clk);
    cudaError_t e = cudaDeviceSynchronize();
    long long h = 0; cudaMemcpy(&h, clk, 8, cudaMemcpyDeviceToHost);
    const double elems = 16.0 * iters * 64 * 32;
    printf("%-58s %9lld clk  %6.2f exps/clk/SM  %5.0f clk per 4-warp round  (%s)\n", name, h, elems / h, (double)h / iters, cudaGetErrorString(e));
    cudaFree(d); cudaFree(clk);
}

int main() {
    run<0>("LDTM + FADD + EX2");
    run<1>("  + F2FP pack");
    run<2>("  + STS.128 swizzled (P tile)");
    run<3>("  + fence.proxy.async + mbarrier.arrive per unit");
    run<4>("  + sampled max (4 FMNMX per piece)");
    return 0;
}
