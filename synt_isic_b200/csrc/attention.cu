// Low-resolution self-attention core of the UNet (diffusers Attention + AttnProcessor2_0,
// reached from core/generator/image_generator.py:400):  o = softmax(q k^T / sqrt(d)) v with
// d = 8, heads = C/8, sequence N = H*W (1024 at 32x32, 256 at 16x16).
//
// attention_simt: fp32 online-softmax kernel (verification mode; also bf16 storage).
//   qkv token layout [B, N, 3C] = (q | k | v), head h occupies channels [8h, 8h+8).
//   One block = 128 queries of one (b, head); K/V streamed through shared memory in fp32.
#include "kernels.cuh"

namespace synt {

constexpr int AT_Q = 128, AT_KCHUNK = 256, AT_D = 8;

template <typename T>
__global__ void __launch_bounds__(AT_Q) attention_simt_kernel(const T* __restrict__ qkv, int N, int C,
                                                              T* __restrict__ out) {
    __shared__ float Ks[AT_KCHUNK][AT_D];
    __shared__ float Vs[AT_KCHUNK][AT_D];
    const int head = blockIdx.y, b = blockIdx.z;
    const int qi = blockIdx.x * AT_Q + threadIdx.x;
    const size_t row = (size_t)3 * C;
    const T* base = qkv + (size_t)b * N * row;
    const float scale = 0.35355339059327373f;              // 1/sqrt(8)
    float q[AT_D];
    if (qi < N) {
        float t[8];
        load8<T>(base + (size_t)qi * row + head * AT_D, t);
#pragma unroll
        for (int i = 0; i < AT_D; ++i) q[i] = t[i] * scale;
    } else {
#pragma unroll
        for (int i = 0; i < AT_D; ++i) q[i] = 0.f;
    }
    float m = -INFINITY, l = 0.f, acc[AT_D];
#pragma unroll
    for (int i = 0; i < AT_D; ++i) acc[i] = 0.f;

    for (int k0 = 0; k0 < N; k0 += AT_KCHUNK) {
        __syncthreads();
        for (int i = threadIdx.x; i < AT_KCHUNK; i += AT_Q) {
            float t[8];
            if (k0 + i < N) {
                load8<T>(base + (size_t)(k0 + i) * row + C + head * AT_D, t);
#pragma unroll
                for (int j = 0; j < AT_D; ++j) Ks[i][j] = t[j];
                load8<T>(base + (size_t)(k0 + i) * row + 2 * C + head * AT_D, t);
#pragma unroll
                for (int j = 0; j < AT_D; ++j) Vs[i][j] = t[j];
            }
        }
        __syncthreads();
        const int kn = min(AT_KCHUNK, N - k0);
        for (int kk = 0; kk < kn; kk += 8) {            // N is a multiple of 8
            float sc[8], gmax = m;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                float d = 0.f;
#pragma unroll
                for (int i = 0; i < AT_D; ++i) d = fmaf(q[i], Ks[kk + j][i], d);
                sc[j] = d;
                gmax = fmaxf(gmax, d);
            }
            const float alpha = expf(m - gmax);          // exp(-inf) = 0 on the first group
            m = gmax;
            l *= alpha;
#pragma unroll
            for (int i = 0; i < AT_D; ++i) acc[i] *= alpha;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const float p = expf(sc[j] - gmax);
                l += p;
#pragma unroll
                for (int i = 0; i < AT_D; ++i) acc[i] = fmaf(p, Vs[kk + j][i], acc[i]);
            }
        }
    }
    if (qi < N) {
        const float inv = 1.0f / l;
        float o[8];
#pragma unroll
        for (int i = 0; i < AT_D; ++i) o[i] = acc[i] * inv;
        store8<T>(out + ((size_t)b * N + qi) * C + head * AT_D, o);
    }
}

void attention_simt(const void* qkv, int dt, int B, int N, int C, void* out, cudaStream_t s) {
    SYNT_CHECK(C % 8 == 0 && N % 8 == 0, "attention: C, N must be multiples of 8");
    dim3 grid(ceil_div(N, AT_Q), C / AT_D, B);
    if (dt == DT_F32) attention_simt_kernel<float><<<grid, AT_Q, 0, s>>>((const float*)qkv, N, C, (float*)out);
    else              attention_simt_kernel<bf16><<<grid, AT_Q, 0, s>>>((const bf16*)qkv, N, C, (bf16*)out);
    SYNT_LAUNCH_CHECK();
}

}  // namespace synt
