// fp32-FMA implicit-GEMM convolution: the fp32 VERIFICATION mode of the UNet/ResNet path
// (north_star: "<=1e-5 in an fp32 verification mode") and the carrier for shapes the
// tcgen05 kernel does not take (Cin=3 stem).  Same ConvArgs contract as conv_tc.
//
// GEMM view: M = B*Ho*Wo pixels, N = Cout, K = KH*KW*Cin (+ 1x1 shortcut segments).
// Tile 64x64x16, 256 threads, 4x4 accumulators per thread, fp32 weights [Cout][Ktot].
#include "kernels.cuh"

namespace synt {

constexpr int SB_M = 64, SB_N = 64, SB_K = 16;

struct SimtGeom {
    int B, H, W, Cin, KH, KW, stride, pad, Ho, Wo, Cout;
    int sc0_C, sc1_C, sc_stride, Kmain, Ktot;
    long long M;
};

template <typename T> __device__ __forceinline__ void load4(const T* p, float (&v)[4]);
template <> __device__ __forceinline__ void load4<float>(const float* p, float (&v)[4]) {
    const float4 a = *reinterpret_cast<const float4*>(p);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
}
template <> __device__ __forceinline__ void load4<bf16>(const bf16* p, float (&v)[4]) {
    const uint2 r = *reinterpret_cast<const uint2*>(p);
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&r);
    const float2 a = __bfloat1622float2(h[0]), b = __bfloat1622float2(h[1]);
    v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y;
}

// element k of the im2col row of output pixel (b, oy, ox); zero outside the image / K range
template <typename T>
__device__ __forceinline__ const T* a_ptr(const SimtGeom& g, const T* in, const T* sc0, const T* sc1, long long b,
                                          int oy, int ox, int k) {
    if (k < g.Kmain) {
        const int tap = k / g.Cin, c = k - tap * g.Cin;
        const int dy = tap / g.KW, dx = tap - dy * g.KW;
        const int iy = oy * g.stride - g.pad + dy, ix = ox * g.stride - g.pad + dx;
        if (iy < 0 || iy >= g.H || ix < 0 || ix >= g.W) return nullptr;
        return in + ((b * g.H + iy) * g.W + ix) * g.Cin + c;
    }
    if (k >= g.Ktot) return nullptr;
    int kk = k - g.Kmain;
    const long long pix = (b * (g.Ho * g.sc_stride) + oy * g.sc_stride) * (long long)(g.Wo * g.sc_stride) + ox * g.sc_stride;
    if (kk < g.sc0_C) return sc0 + pix * g.sc0_C + kk;
    kk -= g.sc0_C;
    return sc1 + pix * g.sc1_C + kk;
}

template <typename T, bool VEC>
__global__ void __launch_bounds__(256) conv_simt_kernel(const SimtGeom g, const T* __restrict__ in,
                                                        const T* __restrict__ sc0, const T* __restrict__ sc1,
                                                        const float* __restrict__ wt, const float* __restrict__ bias,
                                                        const float* __restrict__ bias2, const T* __restrict__ residual,
                                                        int relu, T* __restrict__ out) {
    __shared__ float As[SB_K][SB_M + 4];
    __shared__ float Bs[SB_K][SB_N + 4];
    const int tid = threadIdx.x;
    const long long m0 = (long long)blockIdx.x * SB_M;
    const int n0 = blockIdx.y * SB_N;
    const int lrow = tid >> 2, kq = (tid & 3) * 4;
    // A-row owned by this thread for loading
    const long long lm = m0 + lrow;
    const bool lm_ok = lm < g.M;
    int lox = 0, loy = 0; long long lb = 0;
    if (lm_ok) { lox = (int)(lm % g.Wo); loy = (int)((lm / g.Wo) % g.Ho); lb = lm / ((long long)g.Wo * g.Ho); }
    const int ln = n0 + lrow;
    const int ty = tid >> 4, tx = tid & 15;
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    for (int k0 = 0; k0 < g.Ktot; k0 += SB_K) {
        float a[4] = {0.f, 0.f, 0.f, 0.f}, bw[4] = {0.f, 0.f, 0.f, 0.f};
        const int kk = k0 + kq;
        if (VEC) {
            if (lm_ok) { const T* p = a_ptr<T>(g, in, sc0, sc1, lb, loy, lox, kk); if (p) load4<T>(p, a); }
            if (ln < g.Cout) load4<float>(wt + (size_t)ln * g.Ktot + kk, bw);
        } else {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                if (lm_ok) { const T* p = a_ptr<T>(g, in, sc0, sc1, lb, loy, lox, kk + j); if (p) a[j] = to_f<T>(*p); }
                if (ln < g.Cout && kk + j < g.Ktot) bw[j] = wt[(size_t)ln * g.Ktot + kk + j];
            }
        }
        __syncthreads();      // previous chunk fully consumed
#pragma unroll
        for (int j = 0; j < 4; ++j) { As[kq + j][lrow] = a[j]; Bs[kq + j][lrow] = bw[j]; }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < SB_K; ++k) {
            const float4 av = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
            const float4 bv = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
            const float am[4] = {av.x, av.y, av.z, av.w}, bn[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(am[i], bn[j], acc[i][j]);
        }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const long long m = m0 + ty * 4 + i;
        if (m >= g.M) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int n = n0 + tx * 4 + j;
            if (n >= g.Cout) continue;
            float v = acc[i][j] + bias[n];
            if (bias2) v += bias2[n];
            if (residual) v += to_f<T>(residual[m * g.Cout + n]);
            if (relu) v = fmaxf(v, 0.f);
            out[m * g.Cout + n] = from_f<T>(v);
        }
    }
}

void conv_simt(const ConvArgs& a, int dt, cudaStream_t s) {
    SimtGeom g;
    g.B = a.B; g.H = a.H; g.W = a.W; g.Cin = a.Cin; g.KH = a.KH; g.KW = a.KW; g.stride = a.stride; g.pad = a.pad;
    g.Ho = a.Ho; g.Wo = a.Wo; g.Cout = a.Cout; g.sc0_C = a.sc0_C; g.sc1_C = a.sc1_C; g.sc_stride = a.sc_stride;
    g.Kmain = a.KH * a.KW * a.Cin; g.Ktot = a.ktot(); g.M = (long long)a.B * a.Ho * a.Wo;
    SYNT_CHECK(a.bias != nullptr, "conv_simt: bias required");
    SYNT_CHECK(a.Cin1 == 0 && a.gn_mode == 0, "conv_simt: concat / fused-GroupNorm inputs are conv_tc2 features");
    const bool vec = (a.Cin % 16 == 0) && (a.sc0_C % 16 == 0) && (a.sc1_C % 16 == 0);
    dim3 grid((unsigned)((g.M + SB_M - 1) / SB_M), ceil_div(a.Cout, SB_N));
#define GO(T, V)                                                                                                 \
    conv_simt_kernel<T, V><<<grid, 256, 0, s>>>(g, (const T*)a.in, (const T*)a.sc0, (const T*)a.sc1,             \
                                                (const float*)a.weight, a.bias, a.bias2, (const T*)a.residual,  \
                                                a.relu, (T*)a.out)
    if (dt == DT_F32) { if (vec) GO(float, true); else GO(float, false); }
    else              { if (vec) GO(bf16, true);  else GO(bf16, false); }
#undef GO
    SYNT_LAUNCH_CHECK();
}

}  // namespace synt
