// tcgen05.mma with the A operand in TENSOR MEMORY (M = 128, bf16): layout check before the attention kernel relies on it.
// A[128][64] is written by four warps with tcgen05.st.32x32b (lane = row, 32-bit column c = elements 2c, 2c+1 of the row), B is
// a [16][64] K-major SWIZZLE_128B tile in shared memory, D[128][16] fp32 accumulates four K steps and is read back.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I../../synt_isic_b200/csrc -o bin/mma_ts mma_ts.cu && bin/mma_ts
#include "../../synt_isic_b200/csrc/ptx.cuh"
#include <cuda_bf16.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cmath>
using namespace synt::ptx;

__device__ __forceinline__ void umma_bf16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}

__global__ void __launch_bounds__(128) k(const __nv_bfloat16* A, const __nv_bfloat16* B, float* D) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, r = threadIdx.x;
    if (warp == 0) tmem_alloc<128>(&slot);
    if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
    // B tile: row n (16 rows) x 64 k, swizzled
    for (int i = threadIdx.x; i < 16 * 64; i += 128) {
        const int n = i >> 6, kk = i & 63;
        *reinterpret_cast<__nv_bfloat16*>(smem + n * 128 + ((((kk >> 3) ^ (n & 7)) << 4) | ((kk & 7) * 2))) = B[n * 64 + kk];
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = slot;
    // A row r -> TMEM lanes, columns 32..63 (32 packed columns = 64 bf16)
    uint32_t v[32];
    for (int c = 0; c < 32; ++c) {
        const uint32_t lo = __bfloat16_as_ushort(A[r * 64 + 2 * c]), hi = __bfloat16_as_ushort(A[r * 64 + 2 * c + 1]);
        v[c] = lo | (hi << 16);
    }
    const uint32_t lane_addr = tmem + ((uint32_t)(warp * 32) << 16);
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
        ::"r"(lane_addr + 32), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]),
          "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]), "r"(v[17]), "r"(v[18]), "r"(v[19]), "r"(v[20]),
          "r"(v[21]), "r"(v[22]), "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]), "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31])
        : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    tc_fence_before();
    __syncthreads();
    if (warp == 0 && elect_one()) {
        tc_fence_after();
        constexpr uint32_t idesc = make_idesc_bf16(128, 16);
        const uint64_t db = make_smem_desc_sw128(smem_u32(smem));
        for (int kk = 0; kk < 4; ++kk) umma_bf16_ts(tmem + 0, tmem + 32 + 8 * kk, db + 2 * kk, idesc, kk != 0);
        umma_commit(&bar);
    }
    mbar_wait(&bar, 0);
    tc_fence_after();
    uint32_t o[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(o[0]), "=r"(o[1]), "=r"(o[2]), "=r"(o[3]), "=r"(o[4]), "=r"(o[5]), "=r"(o[6]), "=r"(o[7]), "=r"(o[8]), "=r"(o[9]),
          "=r"(o[10]), "=r"(o[11]), "=r"(o[12]), "=r"(o[13]), "=r"(o[14]), "=r"(o[15])
        : "r"(lane_addr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int n = 0; n < 16; ++n) D[r * 16 + n] = __uint_as_float(o[n]);
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc<128>(tmem);
}

int main() {
    std::vector<__nv_bfloat16> hA(128 * 64), hB(16 * 64);
    std::vector<float> fA(128 * 64), fB(16 * 64);
    uint32_t st = 1;
    auto rnd = [&]() { st = st * 1664525u + 1013904223u; return ((st >> 8) & 0xffff) / 65536.0f - 0.5f; };
    for (size_t i = 0; i < hA.size(); ++i) { hA[i] = __float2bfloat16(rnd()); fA[i] = __bfloat162float(hA[i]); }
    for (size_t i = 0; i < hB.size(); ++i) { hB[i] = __float2bfloat16(rnd()); fB[i] = __bfloat162float(hB[i]); }
    __nv_bfloat16 *dA, *dB; float* dD;
    cudaMalloc(&dA, hA.size() * 2); cudaMalloc(&dB, hB.size() * 2); cudaMalloc(&dD, 128 * 16 * 4);
    cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice);
    cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice);
    k<<<1, 128, 4096>>>(dA, dB, dD);
    cudaError_t e = cudaDeviceSynchronize();
    std::vector<float> hD(128 * 16);
    cudaMemcpy(hD.data(), dD, hD.size() * 4, cudaMemcpyDeviceToHost);
    double maxerr = 0;
    for (int r = 0; r < 128; ++r)
        for (int n = 0; n < 16; ++n) {
            double ref = 0;
            for (int kk = 0; kk < 64; ++kk) ref += (double)fA[r * 64 + kk] * fB[n * 64 + kk];
            maxerr = fmax(maxerr, fabs(ref - hD[r * 16 + n]));
        }
    printf("tcgen05.mma A-in-TMEM: %s, max |D - ref| = %.3e (D[0][0..3] = %f %f %f %f)\n", cudaGetErrorString(e), maxerr, hD[0], hD[1], hD[2], hD[3]);
    return maxerr < 1e-3 ? 0 : 1;
}
