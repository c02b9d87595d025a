// Issue-slot budget next to the MUFU pipe (sm_100a): how many FMA- / ALU-pipe instructions fit beside one MUFU.EX2 per
// element before the exponential stream slows down, and what the packed f32x2 forms cost.  8 independent chains per thread.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o issue_mix issue_mix.cu && ./issue_mix
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

// OP: 0 ex2 only; 1..8: ex2 + OP fadd; 10 fma.f32x2 only; 11 add.f32x2 only; 12 fadd only; 13 F2FP only; 14 fmnmx only; 15 imad only
// 20+n: ex2 + n x fma.f32x2 (on pairs); 30+n: ex2 + n x (fmnmx)  40+n: ex2 + n imad ; 50+n ex2 + n F2FP
template <int OP>
__global__ void k(uint32_t* out, int iters) {
    float a[8], b[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { a[i] = 0.001f * (threadIdx.x * 8 + i); b[i] = 1.0f + 0.01f * i; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const bool mufu = OP < 10 || OP >= 20;
            if (mufu) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
            if (OP >= 1 && OP <= 8) {
#pragma unroll
                for (int j = 0; j < OP; ++j) asm volatile("add.rn.ftz.f32 %0, %0, %1;" : "+f"(b[i]) : "f"(1.5f));
            }
            if (OP == 12) asm volatile("add.rn.ftz.f32 %0, %0, %1;" : "+f"(b[i]) : "f"(1.5f));
            if (OP == 14) asm volatile("max.ftz.f32 %0, %0, %1;" : "+f"(b[i]) : "f"(a[(i + 1) & 7]));
            if (OP == 15) asm volatile("mad.lo.s32 %0, %0, %1, %0;" : "+r"(*(int*)&b[i]) : "r"(8388608));
            if (OP == 13) asm volatile("cvt.rn.bf16x2.f32 %0, %0, %1;" : "+r"(*(uint32_t*)&b[i]) : "f"(a[i]));
            if ((OP == 10 || OP == 11) && (i & 1) == 0) {
                if (OP == 10) asm volatile("{.reg .b64 x, y; mov.b64 x, {%0, %1}; mov.b64 y, {%2, %3}; fma.rn.ftz.f32x2 x, x, y, y; mov.b64 {%0, %1}, x;}"
                                           : "+f"(a[i]), "+f"(a[i + 1]) : "f"(b[i]), "f"(b[i + 1]));
                else asm volatile("{.reg .b64 x, y; mov.b64 x, {%0, %1}; mov.b64 y, {%2, %3}; add.rn.ftz.f32x2 x, x, y; mov.b64 {%0, %1}, x;}"
                                  : "+f"(a[i]), "+f"(a[i + 1]) : "f"(b[i]), "f"(b[i + 1]));
            }
            if (OP >= 20 && OP < 30 && (i & 1) == 0) {
#pragma unroll
                for (int j = 0; j < OP - 20; ++j)
                    asm volatile("{.reg .b64 x, y; mov.b64 x, {%0, %1}; mov.b64 y, {%2, %3}; fma.rn.ftz.f32x2 x, x, y, y; mov.b64 {%0, %1}, x;}"
                                 : "+f"(b[i]), "+f"(b[i + 1]) : "f"(1.0001f), "f"(1.0002f));
            }
            if (OP >= 30 && OP < 40) {
#pragma unroll
                for (int j = 0; j < OP - 30; ++j) asm volatile("max.ftz.f32 %0, %0, %1;" : "+f"(b[i]) : "f"(-126.0f + j));
            }
            if (OP >= 40 && OP < 50) {
#pragma unroll
                for (int j = 0; j < OP - 40; ++j) asm volatile("mad.lo.s32 %0, %0, %1, %0;" : "+r"(*(int*)&b[i]) : "r"(8388608));
            }
            if (OP >= 60 && OP < 70) {
#pragma unroll
                for (int j = 0; j < OP - 60; ++j) asm volatile("fma.rn.ftz.f32 %0, %0, %1, %2;" : "+f"(b[i]) : "f"(1.0001f), "f"(1.5f));
            }
            if (OP >= 70 && OP < 80 && (i & 1) == 0) {
#pragma unroll
                for (int j = 0; j < OP - 70; ++j)
                    asm volatile("{.reg .b64 x, y; mov.b64 x, {%0, %1}; mov.b64 y, {%2, %3}; add.rn.ftz.f32x2 x, x, y; mov.b64 {%0, %1}, x;}"
                                 : "+f"(b[i]), "+f"(b[i + 1]) : "f"(1.0001f), "f"(1.0002f));
            }
            if (OP >= 80 && OP < 90) {
#pragma unroll
                for (int j = 0; j < OP - 80; ++j) asm volatile("mul.rn.ftz.f32 %0, %0, %1;" : "+f"(b[i]) : "f"(1.0001f));
            }
            if (OP >= 90 && OP < 100) {
#pragma unroll
                for (int j = 0; j < OP - 90; ++j) asm volatile("add.rn.ftz.f32 %0, %0, %1;" : "+f"(b[i]) : "f"(a[(i + 3) & 7]));
            }
            if (OP >= 50 && OP < 60) {
#pragma unroll
                for (int j = 0; j < OP - 50; ++j) asm volatile("cvt.rn.bf16x2.f32 %0, %0, %1;" : "+r"(*(uint32_t*)&b[i]) : "f"(1.25f));
            }
        }
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += a[i] + b[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = __float_as_uint(s);
}

template <int OP>
void run(const char* name, int threads, int blocks_per_sm) {
    uint32_t* d; cudaMalloc(&d, 148 * 8 * 1024 * 4);
    const int iters = 2048, blocks = 148 * blocks_per_sm;
    k<OP><<<blocks, threads>>>(d, 16);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k<OP><<<blocks, threads>>>(d, iters);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    int clk_khz; cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
    const double slots = (double)blocks * threads * iters * 8;          // one per chain step
    const double per_clk_sm = slots / (ms * 1e-3) / (clk_khz * 1e3) / 148.0;
    printf("%-34s warps/SMSP %2d  %8.3f ms  %7.2f chain-steps/clk/SM   clk per warp-step per SMSP %6.2f\n", name,
           threads * blocks_per_sm / 128, ms, per_clk_sm, 128.0 / per_clk_sm);
    cudaFree(d);
}
#define R(OP, name) run<OP>(name, 512, 1); run<OP>(name, 1024, 2);
int main() {
    R(0, "ex2 only");
    R(63, "ex2 + 3 ffma"); R(64, "ex2 + 4 ffma"); R(66, "ex2 + 6 ffma"); R(68, "ex2 + 8 ffma");
    R(74, "ex2 + 2 add.f32x2 per element"); R(76, "ex2 + 3 add.f32x2 per element"); R(78, "ex2 + 4 add.f32x2 per element");
    R(84, "ex2 + 4 fmul"); R(94, "ex2 + 4 fadd (register operand)"); R(98, "ex2 + 8 fadd (register operand)");
    return 0;
    R(1, "ex2 + 1 fadd"); R(2, "ex2 + 2 fadd"); R(3, "ex2 + 3 fadd"); R(4, "ex2 + 4 fadd"); R(6, "ex2 + 6 fadd"); R(8, "ex2 + 8 fadd");
    R(12, "fadd only"); R(10, "fma.f32x2 only (per pair: /2)"); R(11, "add.f32x2 only (per pair: /2)");
    R(13, "F2FP only"); R(14, "fmnmx only"); R(15, "imad only");
    R(22, "ex2 + 1 fma2 per element"); R(24, "ex2 + 2 fma2 per element"); R(26, "ex2 + 3 fma2 per element"); R(28, "ex2 + 4 fma2 per element");
    R(32, "ex2 + 2 fmnmx"); R(34, "ex2 + 4 fmnmx"); R(42, "ex2 + 2 imad"); R(44, "ex2 + 4 imad"); R(51, "ex2 + 1 F2FP"); R(52, "ex2 + 2 F2FP");
    return 0;
}
