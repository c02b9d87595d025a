// HBM-bound kernels of the UNet path: GroupNorm (stats / finalize / apply+SiLU with
// channel concat), nearest 2x upsample, the Cin=3 / Cout=3 edge convolutions (conv_out
// fused with GroupNorm+SiLU on load and DDPMScheduler.step in the epilogue), the
// stand-alone scheduler step, time-embedding tables and format helpers.
//
// Reference semantics (restated, not copied):
//   diffusers ResnetBlock2D / GroupNorm / Upsample2D  <- core/generator/image_generator.py:400
//   diffusers DDPMScheduler.step                      <- core/generator/image_generator.py:403
//   output conversion                                 <- core/generator/image_generator.py:441-447
#include "kernels.cuh"
#include <cstdlib>

namespace synt {

#ifndef SYNT_PDL_DEFAULT
#define SYNT_PDL_DEFAULT 2
#endif
int pdl_mode() {
    static int mode = -1;
    // measured (B=64, 133 kernel nodes per step, round 2): 8.914 ms/step without the attribute, 9.030 with it on every kernel,
    // 8.878 with it on the GEMM kernels only (default)
    if (mode < 0) { const char* e = getenv("SYNT_PDL"); mode = e ? atoi(e) : SYNT_PDL_DEFAULT; }
    return mode;
}

// =============================================================== GroupNorm ==========
int gn_num_chunks(int B, int HW) {
    int want = ceil_div(148 * 6, B);
    int maxc = HW / 32 > 0 ? HW / 32 : 1;
    int n = want < maxc ? want : maxc;
    return n < 1 ? 1 : n;
}

template <typename T>
__global__ void __launch_bounds__(256) gn_stats_kernel(const T* __restrict__ src0, int C0,
                                                       const T* __restrict__ src1, int C1, int HW, int G,
                                                       float2* __restrict__ partials, int nchunk) {
    pdl_enter();
    __shared__ float sm_s[2048];
    __shared__ float sm_q[2048];
    const int C = C0 + C1, nvec = C >> 3, lanes = 256 / nvec;
    const int b = blockIdx.y, chunk = blockIdx.x;
    const int ppc = (HW + nchunk - 1) / nchunk;
    const int p0 = chunk * ppc, p1 = min(HW, p0 + ppc);
    const int v = threadIdx.x % nvec, lane = threadIdx.x / nvec;
    float s[8], q[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { s[i] = 0.f; q[i] = 0.f; }
    if (lane < lanes) {
        const bool first = v * 8 < C0;
        const T* base = first ? src0 + (size_t)b * HW * C0 + v * 8 : src1 + (size_t)b * HW * C1 + (v * 8 - C0);
        const int Cs = first ? C0 : C1;
        int p = p0 + lane;
        for (; p + 3 * lanes < p1; p += 4 * lanes) {          // 4 independent 16/32-byte loads in flight
            float x[4][8];
#pragma unroll
            for (int u = 0; u < 4; ++u) load8<T>(base + (size_t)(p + u * lanes) * Cs, x[u]);
#pragma unroll
            for (int u = 0; u < 4; ++u)
#pragma unroll
                for (int i = 0; i < 8; ++i) { s[i] += x[u][i]; q[i] = fmaf(x[u][i], x[u][i], q[i]); }
        }
        for (; p < p1; p += lanes) {
            float x[8];
            load8<T>(base + (size_t)p * Cs, x);
#pragma unroll
            for (int i = 0; i < 8; ++i) { s[i] += x[i]; q[i] = fmaf(x[i], x[i], q[i]); }
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) { sm_s[lane * C + v * 8 + i] = s[i]; sm_q[lane * C + v * 8 + i] = q[i]; }
    }
    __syncthreads();
    if (threadIdx.x < G) {
        const int cpg = C / G, g = threadIdx.x;
        float ts = 0.f, tq = 0.f;
        for (int l = 0; l < lanes; ++l)
            for (int c = 0; c < cpg; ++c) { ts += sm_s[l * C + g * cpg + c]; tq += sm_q[l * C + g * cpg + c]; }
        partials[((size_t)b * nchunk + chunk) * G + g] = make_float2(ts, tq);
    }
}

void gn_stats(const void* src0, int C0, const void* src1, int C1, int dt, int B, int HW, int G, float2* partials,
              int nchunk, cudaStream_t s) {
    const int C = C0 + C1;
    SYNT_CHECK(C % 8 == 0 && C0 % 8 == 0 && C <= 512 && C % G == 0 && G <= 256, "gn_stats: bad channel counts");
    dim3 grid(nchunk, B);
    if (dt == DT_F32)
        launch_pdl(gn_stats_kernel<float>, grid, dim3(256), 0, s, (const float*)src0, C0, (const float*)src1, C1, HW, G, partials, nchunk);
    else
        launch_pdl(gn_stats_kernel<bf16>, grid, dim3(256), 0, s, (const bf16*)src0, C0, (const bf16*)src1, C1, HW, G, partials, nchunk);
    SYNT_LAUNCH_CHECK();
}

// one block per image: 8 threads per group sum the chunk partials (fixed order -> deterministic),
// then every channel gets scale = rstd*gamma, shift = beta - mean*scale
__global__ void __launch_bounds__(256) gn_finalize_kernel(const float2* __restrict__ partials, int nchunk, int G, int C,
                                                          int HW, float eps, const float* __restrict__ gamma,
                                                          const float* __restrict__ beta, float2* __restrict__ scale_shift) {
    __shared__ float2 stat[32];                              // (mean, rstd) per group, G <= 32
    const int b = blockIdx.x, g = threadIdx.x >> 3, part = threadIdx.x & 7;
    if (g < G) {
        double ts = 0.0, tq = 0.0;
        for (int k = part; k < nchunk; k += 8) {
            const float2 p = partials[((size_t)b * nchunk + k) * G + g];
            ts += (double)p.x; tq += (double)p.y;
        }
#pragma unroll
        for (int o = 4; o > 0; o >>= 1) {
            ts += __shfl_xor_sync(0xffffffffu, ts, o);
            tq += __shfl_xor_sync(0xffffffffu, tq, o);
        }
        if (part == 0) {
            const double cnt = (double)HW * (C / G);
            const double mean = ts / cnt;
            double var = tq / cnt - mean * mean;
            if (var < 0.0) var = 0.0;
            stat[g] = make_float2((float)mean, (float)(1.0 / sqrt(var + (double)eps)));
        }
    }
    __syncthreads();
    const int cpg = C / G;
    for (int c = threadIdx.x; c < C; c += 256) {
        const float2 st = stat[c / cpg];
        const float sc = st.y * gamma[c];
        scale_shift[(size_t)b * C + c] = make_float2(sc, beta[c] - st.x * sc);
    }
}

void gn_finalize(const float2* partials, int B, int nchunk, int G, int C, int HW, float eps, const float* gamma,
                 const float* beta, float2* scale_shift, cudaStream_t s) {
    SYNT_CHECK(C <= 512 && G <= 32, "gn_finalize: C <= 512, G <= 32");
    gn_finalize_kernel<<<B, 256, 0, s>>>(partials, nchunk, G, C, HW, eps, gamma, beta, scale_shift);
    SYNT_LAUNCH_CHECK();
}

// Per-channel partials -> scale/shift.  One block per image; channel sums in double, fixed order.
__global__ void __launch_bounds__(256) gn_finalize_channels_kernel(const float2* __restrict__ partA, int SA, int C0,
                                                                   const float2* __restrict__ partB, int SB, int C1, int G,
                                                                   int HW, float eps, const float* __restrict__ gamma,
                                                                   const float* __restrict__ beta,
                                                                   float2* __restrict__ scale_shift) {
    pdl_enter();
    __shared__ double cs[4][512];
    __shared__ double cq[4][512];
    __shared__ float2 stat[32];
    const int b = blockIdx.x, C = C0 + C1;
    const int nslice = C <= 256 ? 256 / C : 1;              // threads cooperating on one channel
    auto channel_sum = [&](int c, int slice) {
        const bool first = c < C0;
        const float2* base = first ? partA + (size_t)b * SA * C0 + c : partB + (size_t)b * SB * C1 + (c - C0);
        const int S = first ? SA : SB, Cs = first ? C0 : C1;
        double ts = 0.0, tq = 0.0;
        int k = slice;
        for (; k + 7 * nslice < S; k += 8 * nslice) {          // 8 independent loads in flight (the chain was latency bound)
            float2 v[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] = __ldg(base + (size_t)(k + j * nslice) * Cs);
#pragma unroll
            for (int j = 0; j < 8; ++j) { ts += (double)v[j].x; tq += (double)v[j].y; }
        }
        for (; k < S; k += nslice) {
            const float2 v = __ldg(base + (size_t)k * Cs);
            ts += (double)v.x; tq += (double)v.y;
        }
        cs[slice][c] = ts; cq[slice][c] = tq;
    };
    if (C <= 256) {
        const int c = threadIdx.x % C, slice = threadIdx.x / C;
        if (slice < nslice) channel_sum(c, slice);
    } else {
        for (int c = threadIdx.x; c < C; c += 256) channel_sum(c, 0);
    }
    __syncthreads();
    const int cpg = C / G;
    if (threadIdx.x < G) {
        double ts = 0.0, tq = 0.0;
        for (int c = threadIdx.x * cpg; c < (threadIdx.x + 1) * cpg; ++c)
            for (int sl = 0; sl < nslice; ++sl) { ts += cs[sl][c]; tq += cq[sl][c]; }
        const double cnt = (double)HW * cpg;
        const double mean = ts / cnt;
        double var = tq / cnt - mean * mean;
        if (var < 0.0) var = 0.0;
        stat[threadIdx.x] = make_float2((float)mean, (float)(1.0 / sqrt(var + (double)eps)));
    }
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += 256) {
        const float2 st = stat[c / cpg];
        const float sc = st.y * gamma[c];
        scale_shift[(size_t)b * C + c] = make_float2(sc, beta[c] - st.x * sc);
    }
}
void gn_finalize_channels(const float2* partA, int SA, int C0, const float2* partB, int SB, int C1, int B, int G, int HW,
                          float eps, const float* gamma, const float* beta, float2* scale_shift, cudaStream_t s) {
    SYNT_CHECK(C0 + C1 <= 512 && G <= 32 && (C0 + C1) % G == 0, "gn_finalize_channels: bad channel counts");
    launch_pdl(gn_finalize_channels_kernel, dim3(B), dim3(256), 0, s, partA, SA, C0, partB, SB, C1, G, HW, eps, gamma, beta, scale_shift);
}

// silu(x) = x*sigmoid(x) = h + h*tanh(h), h = x/2: one MUFU op (tanh.approx) per element
__device__ __forceinline__ float silu_tanh(float x) {
    const float h = 0.5f * x;
    float t;
    asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(h));
    return fmaf(h, t, h);
}

// GroupNorm (+SiLU) apply with the FINALIZE folded in: every block first reduces the per-channel partial
// sums of its image (a handful of rows per tensor) to scale/shift in shared memory, then each thread
// owns ONE 8-channel vector position (its scale/shift live in registers) and streams the pixels of the
// block's range with 4 independent 16-byte loads in flight.
struct GnSrc { const void* x; const float2* part; int C; int slots; };

template <typename T, bool SILU, bool FUSED>
__global__ void __launch_bounds__(256, 4) gn_apply_kernel(GnSrc a, GnSrc b2, int HW, int ppb, int G, float eps,
                                                       const float* __restrict__ gamma, const float* __restrict__ beta,
                                                       const float2* __restrict__ scale_shift, T* __restrict__ out) {
    pdl_enter();
    __shared__ double csum[FUSED ? 512 : 1];
    __shared__ double csq[FUSED ? 512 : 1];
    __shared__ float2 stat[32];
    __shared__ float2 ss_s[FUSED ? 512 : 1];
    const int C0 = a.C, C1 = b2.C, C = C0 + C1, nvec = C >> 3, rows = 256 / nvec;
    const int b = blockIdx.y;
    if (FUSED) {                                             // finalize in-kernel from the partial rows
    for (int c = threadIdx.x; c < C; c += 256) {
        const bool first = c < C0;
        const float2* base = first ? a.part + (size_t)b * a.slots * C0 + c : b2.part + (size_t)b * b2.slots * C1 + (c - C0);
        const int S = first ? a.slots : b2.slots, Cs = first ? C0 : C1;
        double ts = 0.0, tq = 0.0;
        for (int k = 0; k < S; ++k) {
            const float2 v = __ldg(base + (size_t)k * Cs);
            ts += (double)v.x; tq += (double)v.y;
        }
        csum[c] = ts; csq[c] = tq;
    }
    __syncthreads();
    const int cpg = C / G;
    if (threadIdx.x < G) {
        double ts = 0.0, tq = 0.0;
        for (int c = threadIdx.x * cpg; c < (threadIdx.x + 1) * cpg; ++c) { ts += csum[c]; tq += csq[c]; }
        const double cnt = (double)HW * cpg;
        const double mean = ts / cnt;
        double var = tq / cnt - mean * mean;
        if (var < 0.0) var = 0.0;
        stat[threadIdx.x] = make_float2((float)mean, (float)(1.0 / sqrt(var + (double)eps)));
    }
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += 256) {
        const float2 st = stat[c / cpg];
        const float sc = st.y * gamma[c];
        ss_s[c] = make_float2(sc, beta[c] - st.x * sc);
    }
    __syncthreads();
    }
    const int v = threadIdx.x % nvec, prow = threadIdx.x / nvec;
    if (prow >= rows) return;
    const int c = v * 8;
    float sc[8], sh[8];
    if (!FUSED) {                                            // finalize already done by gn_finalize_channels
        const float4* ss = reinterpret_cast<const float4*>(scale_shift + (size_t)b * C + c);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float4 t = __ldg(ss + j);
            sc[2 * j] = t.x; sh[2 * j] = t.y; sc[2 * j + 1] = t.z; sh[2 * j + 1] = t.w;
        }
    } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) { const float2 t = ss_s[c + j]; sc[j] = t.x; sh[j] = t.y; }
    }
    const bool first = c < C0;
    const T* src = first ? (const T*)a.x + (size_t)b * HW * C0 + c : (const T*)b2.x + (size_t)b * HW * C1 + (c - C0);
    const int Cs = first ? C0 : C1;
    T* dst = out + (size_t)b * HW * C + c;
    const int p0 = blockIdx.x * ppb, p1 = min(HW, p0 + ppb);
    constexpr int U = 4;
    for (int p = p0 + prow; p < p1; p += U * rows) {
        float x[U][8];
#pragma unroll
        for (int u = 0; u < U; ++u)
            if (p + u * rows < p1) load8<T>(src + (size_t)(p + u * rows) * Cs, x[u]);
#pragma unroll
        for (int u = 0; u < U; ++u) {
            if (p + u * rows >= p1) break;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                float y = fmaf(x[u][i], sc[i], sh[i]);
                if (SILU) y = (sizeof(T) == 4) ? silu_precise(y) : silu_tanh(y);
                x[u][i] = y;
            }
            store8<T>(dst + (size_t)(p + u * rows) * C, x[u]);
        }
    }
}

void gn_apply_fused(const void* src0, const float2* part0, int slots0, int C0, const void* src1, const float2* part1,
                    int slots1, int C1, int dt, int B, int HW, int G, float eps, const float* gamma, const float* beta,
                    const float2* scale_shift, int silu, void* out, cudaStream_t s) {
    const int C = C0 + C1, nvec = C / 8, rows = 256 / nvec;
    SYNT_CHECK(C <= 512 && G <= 32 && C % G == 0 && C0 % 8 == 0 && C % 8 == 0, "gn_apply_fused: bad channel counts");
    int ppb = rows * (scale_shift ? 16 : 32);                // pixels per thread (more when the finalize prologue runs)
    if (ppb > HW) ppb = HW;
    dim3 grid(ceil_div(HW, ppb), B);
    GnSrc a{src0, part0, C0, slots0}, b{src1, part1, C1, slots1};
#define GO(T, S, F) launch_pdl(gn_apply_kernel<T, S, F>, grid, dim3(256), 0, s, a, b, HW, ppb, G, eps, gamma, beta, scale_shift, (T*)out)
#define GO2(T, S) do { if (scale_shift) GO(T, S, false); else GO(T, S, true); } while (0)
    if (dt == DT_F32) { if (silu) GO2(float, true); else GO2(float, false); }
    else              { if (silu) GO2(bf16, true);  else GO2(bf16, false); }
#undef GO2
#undef GO
    SYNT_LAUNCH_CHECK();
}

// =============================================================== upsample ===========
template <typename T>
__global__ void upsample2x_kernel(const T* __restrict__ in, int H, int W, int C, long long nvec_total, T* __restrict__ out) {
    const int nvec = C >> 3, Wo = 2 * W, Ho = 2 * H;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < nvec_total;
         i += (long long)gridDim.x * blockDim.x) {
        const int v = (int)(i % nvec);
        long long p = i / nvec;
        const int ox = (int)(p % Wo); p /= Wo;
        const int oy = (int)(p % Ho);
        const long long b = p / Ho;
        const T* src = in + ((b * H + (oy >> 1)) * W + (ox >> 1)) * C + v * 8;
        float x[8];
        load8<T>(src, x);
        store8<T>(out + ((b * Ho + oy) * Wo + ox) * C + v * 8, x);
    }
}
void upsample_nearest2x(const void* in, int dt, int B, int H, int W, int C, void* out, cudaStream_t s) {
    SYNT_CHECK(C % 8 == 0, "upsample: C % 8");
    const long long nv = (long long)B * 4 * H * W * (C / 8);
    const int blocks = (int)((nv + 255) / 256 < 148 * 16 ? (nv + 255) / 256 : 148 * 16);
    if (dt == DT_F32) upsample2x_kernel<float><<<blocks, 256, 0, s>>>((const float*)in, H, W, C, nv, (float*)out);
    else              upsample2x_kernel<bf16><<<blocks, 256, 0, s>>>((const bf16*)in, H, W, C, nv, (bf16*)out);
    SYNT_LAUNCH_CHECK();
}

// =============================================================== conv_in (3 -> 64) ==
template <typename T>
__global__ void __launch_bounds__(128) conv_in3_kernel(const float* __restrict__ x, const __grid_constant__ ConvInW w,
                                                       int B, int H, int W, T* __restrict__ out) {
    const long long pix = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (pix >= (long long)B * H * W) return;
    const int ox = (int)(pix % W), oy = (int)((pix / W) % H);
    const long long b = pix / ((long long)H * W);
    float in[27];
#pragma unroll
    for (int dy = 0; dy < 3; ++dy)
#pragma unroll
        for (int dx = 0; dx < 3; ++dx) {
            const int iy = oy + dy - 1, ix = ox + dx - 1;
            const bool ok = iy >= 0 && iy < H && ix >= 0 && ix < W;
#pragma unroll
            for (int c = 0; c < 3; ++c)
                in[(dy * 3 + dx) * 3 + c] = ok ? __ldg(x + ((b * 3 + c) * H + iy) * W + ix) : 0.f;
        }
    T* o = out + pix * 64;
#pragma unroll
    for (int n0 = 0; n0 < 64; n0 += 8) {
        float acc[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] = w.b[n0 + j];
#pragma unroll
        for (int k = 0; k < 27; ++k)
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[j] = fmaf(in[k], w.w[k][n0 + j], acc[j]);
        store8<T>(o + n0, acc);
    }
}
// ---- bf16 production path: register-tiled fp32 FMA (4 pixels x 16 channels per thread, weights broadcast from
// shared memory) with the GroupNorm statistics of the OUTPUT fused (per-channel sum / sum of squares of the stored
// bf16 values, one partial row per 8-row band).  Block = 256 threads = one band of 8 rows x W columns of one image,
// walked in 32-column tiles; warp = (channel quarter cq = warp & 3, row half), lane = pixel group (row, 4 columns): every
// lane of a warp reads the SAME weights (one shared-memory wavefront per LDS.128) and the input rows are padded to a
// stride of 65 floats so that the 32 pixel groups of a warp hit 32 different banks.
constexpr int CI_ROWS = 8, CI_COLS = 32, CI_STRIDE = 65;

__global__ void __launch_bounds__(256, 2) conv_in3_tiled_kernel(const float* __restrict__ x, const __grid_constant__ ConvInW w,
                                                             int H, int W, bf16* __restrict__ out, float2* __restrict__ stats,
                                                             int stats_slots) {
    pdl_enter();
    __shared__ __align__(16) float w_s[27][64];
    __shared__ float in_s[3][CI_ROWS + 2][CI_STRIDE];
    __shared__ __align__(16) float b_s[64];
    __shared__ float2 red_s[2][64];
    const int b = blockIdx.y, y0 = blockIdx.x * CI_ROWS;
    for (int i = threadIdx.x; i < 27 * 64; i += 256) w_s[i / 64][i % 64] = w.w[i / 64][i % 64];
    if (threadIdx.x < 64) b_s[threadIdx.x] = w.b[threadIdx.x];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int cq = warp & 3, row = (warp >> 2) * 4 + (lane >> 3), xg = lane & 7;
    float s1[16], s2[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) { s1[j] = 0.f; s2[j] = 0.f; }
    const float* xb = x + (size_t)b * 3 * H * W;
    // input halo of one tile: 3 x 10 x 34 floats = 1020 elements, 4 per thread; the NEXT tile's elements are fetched into
    // registers before the FMAs of the current tile and stored to shared memory after them (global latency hidden)
    constexpr int NEL = 3 * (CI_ROWS + 2) * (CI_COLS + 2), PER = (NEL + 255) / 256;
    float nxt[PER];
    auto fetch = [&](int x0) {
#pragma unroll
        for (int u = 0; u < PER; ++u) {
            const int i = threadIdx.x + u * 256;
            const int c = i / ((CI_ROWS + 2) * (CI_COLS + 2)), rem = i % ((CI_ROWS + 2) * (CI_COLS + 2));
            const int hy = rem / (CI_COLS + 2), hx = rem % (CI_COLS + 2);
            const int iy = y0 + hy - 1, ix = x0 + hx - 1;
            nxt[u] = (i < NEL && iy >= 0 && iy < H && ix >= 0 && ix < W) ? __ldg(xb + ((size_t)c * H + iy) * W + ix) : 0.f;
        }
    };
    auto commit = [&]() {
#pragma unroll
        for (int u = 0; u < PER; ++u) {
            const int i = threadIdx.x + u * 256;
            if (i < NEL) {
                const int c = i / ((CI_ROWS + 2) * (CI_COLS + 2)), rem = i % ((CI_ROWS + 2) * (CI_COLS + 2));
                in_s[c][rem / (CI_COLS + 2)][rem % (CI_COLS + 2)] = nxt[u];
            }
        }
    };
    fetch(0);
    for (int x0 = 0; x0 < W; x0 += CI_COLS) {
        __syncthreads();                                             // previous tile fully consumed (and w_s visible)
        commit();
        __syncthreads();
        if (x0 + CI_COLS < W) fetch(x0 + CI_COLS);
        float acc[4][16];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float4 bv = *reinterpret_cast<const float4*>(&b_s[cq * 16 + 4 * j]);
#pragma unroll
            for (int p = 0; p < 4; ++p) { acc[p][4 * j] = bv.x; acc[p][4 * j + 1] = bv.y; acc[p][4 * j + 2] = bv.z; acc[p][4 * j + 3] = bv.w; }
        }
#pragma unroll
        for (int dy = 0; dy < 3; ++dy)
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                float v[6];
#pragma unroll
                for (int j = 0; j < 6; ++j) v[j] = in_s[c][row + dy][xg * 4 + j];
#pragma unroll
                for (int dx = 0; dx < 3; ++dx) {
                    const float4* wr = reinterpret_cast<const float4*>(&w_s[(dy * 3 + dx) * 3 + c][cq * 16]);
                    float wv[16];
#pragma unroll
                    for (int j = 0; j < 4; ++j) { const float4 t = wr[j]; wv[4 * j] = t.x; wv[4 * j + 1] = t.y; wv[4 * j + 2] = t.z; wv[4 * j + 3] = t.w; }
#pragma unroll
                    for (int p = 0; p < 4; ++p)
#pragma unroll
                        for (int j = 0; j < 16; ++j) acc[p][j] = fmaf(v[p + dx], wv[j], acc[p][j]);
                }
            }
        const int oy = y0 + row;
#pragma unroll
        for (int p = 0; p < 4; ++p) {
            const int ox = x0 + xg * 4 + p;
            uint4 pk[2];
            __nv_bfloat162* h2 = reinterpret_cast<__nv_bfloat162*>(pk);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                h2[j] = __floats2bfloat162_rn(acc[p][2 * j], acc[p][2 * j + 1]);
                const float2 r = __bfloat1622float2(h2[j]);          // statistics of the values the consumers will read
                s1[2 * j] += r.x; s2[2 * j] = fmaf(r.x, r.x, s2[2 * j]);
                s1[2 * j + 1] += r.y; s2[2 * j + 1] = fmaf(r.y, r.y, s2[2 * j + 1]);
            }
            uint4* o = reinterpret_cast<uint4*>(out + (((size_t)b * H + oy) * W + ox) * 64 + cq * 16);
            o[0] = pk[0]; o[1] = pk[1];
        }
    }
    // per-channel sums over the band: the 32 lanes of a warp, then the two warps that share cq
#pragma unroll
    for (int j = 0; j < 16; ++j) {
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            s1[j] += __shfl_xor_sync(0xffffffffu, s1[j], o);
            s2[j] += __shfl_xor_sync(0xffffffffu, s2[j], o);
        }
    }
    if (lane == 0) {
#pragma unroll
        for (int j = 0; j < 16; ++j) red_s[warp >> 2][cq * 16 + j] = make_float2(s1[j], s2[j]);
    }
    __syncthreads();
    if (threadIdx.x < 64 && stats)
        stats[((size_t)b * stats_slots + blockIdx.x) * 64 + threadIdx.x] =
            make_float2(red_s[0][threadIdx.x].x + red_s[1][threadIdx.x].x, red_s[0][threadIdx.x].y + red_s[1][threadIdx.x].y);
}

int conv_in3_stats_slots(int H, int W, int dt) { return (dt == DT_BF16 && H % CI_ROWS == 0 && W % CI_COLS == 0) ? H / CI_ROWS : 0; }

void conv_in3(const float* x, const ConvInW& w, int B, int H, int W, void* out, int dt, float2* stats_out, cudaStream_t s) {
    if (conv_in3_stats_slots(H, W, dt) > 0) {
        launch_pdl(conv_in3_tiled_kernel, dim3(H / CI_ROWS, B), dim3(256), 0, s, x, w, H, W, (bf16*)out, stats_out, H / CI_ROWS);
        return;
    }
    SYNT_CHECK(stats_out == nullptr, "conv_in3: fused statistics need the tiled bf16 path");
    const long long np = (long long)B * H * W;
    const int blocks = (int)((np + 127) / 128);
    if (dt == DT_F32) conv_in3_kernel<float><<<blocks, 128, 0, s>>>(x, w, B, H, W, (float*)out);
    else              conv_in3_kernel<bf16><<<blocks, 128, 0, s>>>(x, w, B, H, W, (bf16*)out);
    SYNT_LAUNCH_CHECK();
}

// =============================================================== Philox =============
__device__ __forceinline__ void philox4x32_10(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c[0]), lo0 = 0xD2511F53u * c[0];
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c[2]), lo1 = 0xCD9E8D57u * c[2];
        c[0] = hi1 ^ c[1] ^ k0; c[1] = lo1; c[2] = hi0 ^ c[3] ^ k1; c[3] = lo0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
}
__device__ __forceinline__ void philox_normal3(unsigned long long seed, unsigned long long elem, uint32_t step,
                                               float (&z)[3]) {
    uint32_t c[4] = {(uint32_t)elem, (uint32_t)(elem >> 32), step, 0x5eed5eedu};
    philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32));
    const float u0 = ((float)c[0] + 0.5f) * 2.3283064365386963e-10f;   // (0,1)
    const float u1 = ((float)c[1] + 0.5f) * 2.3283064365386963e-10f;
    const float u2 = ((float)c[2] + 0.5f) * 2.3283064365386963e-10f;
    const float u3 = ((float)c[3] + 0.5f) * 2.3283064365386963e-10f;
    const float r0 = sqrtf(-2.f * __logf(u0)), r1 = sqrtf(-2.f * __logf(u2));
    float s0, c0, s1, c1;
    __sincosf(6.283185307179586f * u1, &s0, &c0);
    __sincosf(6.283185307179586f * u3, &s1, &c1);
    z[0] = r0 * c0; z[1] = r0 * s0; z[2] = r1 * c1; (void)s1;
}

// eps of one pixel (3 channels) -> optional eps tap, DDPMScheduler.step on x (in place), optional trajectory frame
__device__ __forceinline__ void sched_epilogue(const float (&e)[3], int b, int oy, int ox, int H, int W, float* eps_out,
                                               const SchedArgs& sch) {
    const size_t plane = (size_t)H * W;
    const size_t i0 = ((size_t)b * 3) * plane + (size_t)oy * W + ox;
    const long long step = sch.step_ptr ? (long long)*sch.step_ptr : 0;
    if (eps_out) {
        float* eo = eps_out + step * sch.eps_step_stride;
#pragma unroll
        for (int j = 0; j < 3; ++j) eo[i0 + j * plane] = e[j];
    }
    if (sch.x) {
        if (sch.step_mask && sch.step_mask[step * sch.mask_stride + b] == 0) {   // this image skips the step: x frozen
            if (sch.traj) {
#pragma unroll
                for (int j = 0; j < 3; ++j) sch.traj[step * sch.traj_step_stride + i0 + j * plane] = sch.x[i0 + j * plane];
            }
            return;
        }
        const float sqrt_b = sch.coef[0], sqrt_a = sch.coef[1], c_x0 = sch.coef[2], c_xt = sch.coef[3],
                    sigma = sch.coef[4];
        float z[3] = {0.f, 0.f, 0.f};
        if (sigma != 0.f) {
            if (sch.z) {
#pragma unroll
                for (int j = 0; j < 3; ++j) z[j] = sch.z[step * sch.z_step_stride + i0 + j * plane];
            } else {
                const unsigned long long img = sch.noise_shared ? (unsigned long long)sch.image_offset
                                             : sch.image_keys ? (unsigned long long)sch.image_keys[b]
                                                              : (unsigned long long)(sch.image_offset + b);
                const unsigned long long elem = img * plane + (size_t)oy * W + ox;
                philox_normal3(sch.seed, elem, (uint32_t)step, z);
            }
        }
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            const float xv = sch.x[i0 + j * plane];
            float x0v = (xv - sqrt_b * e[j]) / sqrt_a;
            x0v = fminf(fmaxf(x0v, -1.f), 1.f);
            float prev = c_x0 * x0v + c_xt * xv;
            if (sigma != 0.f) prev = prev + sigma * z[j];
            sch.x[i0 + j * plane] = prev;
            if (sch.traj) sch.traj[step * sch.traj_step_stride + i0 + j * plane] = prev;
        }
    }
}

// =============================================================== conv_out (64 -> 3) =
// One block = 16x16 output pixels.  The GroupNorm+SiLU'd halo tile (18x18x64) is staged in
// shared memory as fp32 with a padded pixel stride of 65 floats (bank-conflict-free),
// weights come from the constant bank (by-value kernel parameter).
constexpr int CO_TILE = 16, CO_HALO = CO_TILE + 2, CO_STRIDE = 65;
constexpr int CO_SMEM_BYTES = CO_HALO * CO_HALO * CO_STRIDE * 4;

template <typename T>
__global__ void __launch_bounds__(256) conv_out3_kernel(const T* __restrict__ h, const float2* __restrict__ scale_shift,
                                                        const __grid_constant__ ConvOutW w, int B, int H, int W,
                                                        float* __restrict__ eps_out, const __grid_constant__ SchedArgs sch) {
    extern __shared__ float tile[];
    const int b = blockIdx.z, y0 = blockIdx.y * CO_TILE, x0 = blockIdx.x * CO_TILE;
    // ---- stage: 324 halo pixels x 8 vectors of 8 channels
    for (int i = threadIdx.x; i < CO_HALO * CO_HALO * 8; i += 256) {
        const int v = i & 7, hp = i >> 3;
        const int hy = hp / CO_HALO, hx = hp % CO_HALO;
        const int iy = y0 + hy - 1, ix = x0 + hx - 1;
        float x[8];
        if (iy >= 0 && iy < H && ix >= 0 && ix < W) {
            load8<T>(h + (((size_t)b * H + iy) * W + ix) * 64 + v * 8, x);
            const float4* ss = reinterpret_cast<const float4*>(scale_shift + (size_t)b * 64 + v * 8);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                float4 t = __ldg(ss + j);
                float a0 = fmaf(x[2 * j], t.x, t.y), a1 = fmaf(x[2 * j + 1], t.z, t.w);
                if (sizeof(T) == 4) { a0 = silu_precise(a0); a1 = silu_precise(a1); }
                else                { a0 = silu_tanh(a0);    a1 = silu_tanh(a1); }
                x[2 * j] = a0; x[2 * j + 1] = a1;
            }
        } else {
#pragma unroll
            for (int j = 0; j < 8; ++j) x[j] = 0.f;
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) tile[hp * CO_STRIDE + v * 8 + j] = x[j];
    }
    __syncthreads();
    const int ty = threadIdx.x / CO_TILE, tx = threadIdx.x % CO_TILE;
    const int oy = y0 + ty, ox = x0 + tx;
    float acc0 = w.b[0], acc1 = w.b[1], acc2 = w.b[2];
#pragma unroll
    for (int tap = 0; tap < 9; ++tap) {
        const float* src = tile + ((ty + tap / 3) * CO_HALO + tx + tap % 3) * CO_STRIDE;
#pragma unroll 16
        for (int c = 0; c < 64; ++c) {
            const float v = src[c];
            acc0 = fmaf(v, w.w[tap][c][0], acc0);
            acc1 = fmaf(v, w.w[tap][c][1], acc1);
            acc2 = fmaf(v, w.w[tap][c][2], acc2);
        }
    }
    if (oy >= H || ox >= W) return;
    const float e[3] = {acc0, acc1, acc2};
    sched_epilogue(e, b, oy, ox, H, W, eps_out, sch);
}

// ---- bf16 production path: the same tile on the legacy tensor-core path (mma.sync m16n8k16, N = 3 padded to 8).
// The GroupNorm+SiLU'd 18x18x64 halo tile is staged as bf16 with a 16-byte-chunk XOR swizzle (pixel & 7) so that the
// ldmatrix rows of 8 consecutive pixels are bank-conflict free; the weights arrive pre-arranged as per-lane B
// fragments (36 k-steps = 9 taps x 4 sixteen-channel blocks).  Each warp owns two rows of 16 pixels (two m16 tiles).
// eps goes through shared memory so that one thread owns all three channels of a pixel in the scheduler epilogue.
constexpr int COT_TILE_BYTES = CO_HALO * CO_HALO * 128;          // 41472
constexpr int COT_FRAG_BYTES = 36 * 32 * 8;                       // 9216
constexpr int COT_SMEM_BYTES = COT_TILE_BYTES + COT_FRAG_BYTES + 256 * 3 * 4;

__global__ void __launch_bounds__(256) conv_out3_mma_kernel(const bf16* __restrict__ h, const float2* __restrict__ scale_shift,
                                                            const uint2* __restrict__ bfrag, float b0, float b1, float b2,
                                                            int B, int H, int W, float* __restrict__ eps_out,
                                                            const __grid_constant__ SchedArgs sch) {
    pdl_enter();
    extern __shared__ __align__(128) uint8_t cot_smem[];
    uint8_t* tile = cot_smem;
    uint2* frag_s = reinterpret_cast<uint2*>(cot_smem + COT_TILE_BYTES);
    float* eps_s = reinterpret_cast<float*>(cot_smem + COT_TILE_BYTES + COT_FRAG_BYTES);
    const int b = blockIdx.z, y0 = blockIdx.y * CO_TILE, x0 = blockIdx.x * CO_TILE;
    for (int i = threadIdx.x; i < 36 * 32; i += 256) frag_s[i] = __ldg(bfrag + i);
    // ---- stage: 324 halo pixels x 8 vectors of 8 channels, GroupNorm affine + SiLU, bf16
    {
        const int v = threadIdx.x & 7;                               // this thread's channel vector is fixed
        float sc[8], sh[8];
        const float4* ss = reinterpret_cast<const float4*>(scale_shift + (size_t)b * 64 + v * 8);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float4 t = __ldg(ss + j);
            sc[2 * j] = 0.5f * t.x; sh[2 * j] = 0.5f * t.y; sc[2 * j + 1] = 0.5f * t.z; sh[2 * j + 1] = 0.5f * t.w;   // h = y/2
        }
        constexpr int NP = CO_HALO * CO_HALO;
        for (int hp0 = threadIdx.x >> 3; hp0 < NP; hp0 += 4 * 32) {   // 4 pixels in flight per thread
            uint4 raw[4]; bool ok[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int hp = hp0 + u * 32;
                const int hy = hp / CO_HALO, hx = hp - hy * CO_HALO;
                const int iy = y0 + hy - 1, ix = x0 + hx - 1;
                ok[u] = hp < NP && iy >= 0 && iy < H && ix >= 0 && ix < W;
                if (ok[u]) raw[u] = __ldg(reinterpret_cast<const uint4*>(h + (((size_t)b * H + iy) * W + ix) * 64 + v * 8));
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int hp = hp0 + u * 32;
                if (hp >= NP) continue;
                uint4 o = make_uint4(0u, 0u, 0u, 0u);
                if (ok[u]) {
                    const __nv_bfloat162* x2 = reinterpret_cast<const __nv_bfloat162*>(&raw[u]);
                    __nv_bfloat162* o2 = reinterpret_cast<__nv_bfloat162*>(&o);
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const float2 f = __bfloat1622float2(x2[j]);
                        const float h0 = fmaf(f.x, sc[2 * j], sh[2 * j]), h1 = fmaf(f.y, sc[2 * j + 1], sh[2 * j + 1]);
                        float t0, t1;
                        asm("tanh.approx.f32 %0, %1;" : "=f"(t0) : "f"(h0));
                        asm("tanh.approx.f32 %0, %1;" : "=f"(t1) : "f"(h1));
                        o2[j] = __floats2bfloat162_rn(fmaf(h0, t0, h0), fmaf(h1, t1, h1));   // silu(y) = h + h tanh(h)
                    }
                }
                *reinterpret_cast<uint4*>(tile + hp * 128 + ((v ^ (hp & 7)) << 4)) = o;
            }
        }
    }
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = lane >> 2, t = lane & 3;
    float acc[2][4];
#pragma unroll
    for (int m = 0; m < 2; ++m) {
        acc[m][0] = acc[m][2] = t == 0 ? b0 : (t == 1 ? b2 : 0.f);
        acc[m][1] = acc[m][3] = t == 0 ? b1 : 0.f;
    }
    const uint32_t tile_u32 = static_cast<uint32_t>(__cvta_generic_to_shared(tile));
    const int lm = lane >> 3, lr = lane & 7;                         // ldmatrix: matrix index, row within it
#pragma unroll
    for (int tap = 0; tap < 9; ++tap) {
        const int dy = tap / 3, dx = tap % 3;
#pragma unroll
        for (int kc = 0; kc < 4; ++kc) {
            const uint2 bf = frag_s[(tap * 4 + kc) * 32 + lane];
#pragma unroll
            for (int m = 0; m < 2; ++m) {
                const int hp = (warp * 2 + m + dy) * CO_HALO + dx + (lm & 1) * 8 + lr;
                const uint32_t addr = tile_u32 + hp * 128 + (((kc * 2 + (lm >> 1)) ^ (hp & 7)) << 4);
                uint32_t a0, a1, a2, a3;
                asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
                             : "=r"(a0), "=r"(a1), "=r"(a2), "=r"(a3) : "r"(addr));
                asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
                             : "+f"(acc[m][0]), "+f"(acc[m][1]), "+f"(acc[m][2]), "+f"(acc[m][3])
                             : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(bf.x), "r"(bf.y));
            }
        }
    }
    // accumulator fragment -> eps_s[pixel][3]: rows g / g+8 of the m16 tile = pixels tx = g / g+8 of row warp*2+m
#pragma unroll
    for (int m = 0; m < 2; ++m) {
        const int p0 = (warp * 2 + m) * 16 + g;
        if (t == 0) {
            eps_s[p0 * 3 + 0] = acc[m][0]; eps_s[p0 * 3 + 1] = acc[m][1];
            eps_s[(p0 + 8) * 3 + 0] = acc[m][2]; eps_s[(p0 + 8) * 3 + 1] = acc[m][3];
        } else if (t == 1) {
            eps_s[p0 * 3 + 2] = acc[m][0]; eps_s[(p0 + 8) * 3 + 2] = acc[m][2];
        }
    }
    __syncthreads();
    const int ty = threadIdx.x / CO_TILE, tx = threadIdx.x % CO_TILE;
    const int oy = y0 + ty, ox = x0 + tx;
    if (oy >= H || ox >= W) return;
    const float e[3] = {eps_s[threadIdx.x * 3], eps_s[threadIdx.x * 3 + 1], eps_s[threadIdx.x * 3 + 2]};
    sched_epilogue(e, b, oy, ox, H, W, eps_out, sch);
}

void conv_out3(const void* h, int dt, const float2* scale_shift, const ConvOutW& w, const void* bfrag, int B, int H, int W,
               float* eps_nchw, const SchedArgs& sch, cudaStream_t s) {
    dim3 grid(ceil_div(W, CO_TILE), ceil_div(H, CO_TILE), B);
    ensure_dynamic_smem((const void*)(conv_out3_kernel<float>), CO_SMEM_BYTES);
    ensure_dynamic_smem((const void*)(conv_out3_kernel<bf16>), CO_SMEM_BYTES);
    ensure_dynamic_smem((const void*)(conv_out3_mma_kernel), COT_SMEM_BYTES);
    if (dt == DT_F32)
        conv_out3_kernel<float><<<grid, 256, CO_SMEM_BYTES, s>>>((const float*)h, scale_shift, w, B, H, W, eps_nchw, sch);
    else if (bfrag)
        launch_pdl(conv_out3_mma_kernel, grid, dim3(256), COT_SMEM_BYTES, s, (const bf16*)h, scale_shift, (const uint2*)bfrag,
                   w.b[0], w.b[1], w.b[2], B, H, W, eps_nchw, sch);
    else
        conv_out3_kernel<bf16><<<grid, 256, CO_SMEM_BYTES, s>>>((const bf16*)h, scale_shift, w, B, H, W, eps_nchw, sch);
    SYNT_LAUNCH_CHECK();
}

// =============================================================== scheduler step =====
__global__ void ddpm_step_kernel(const float* __restrict__ eps, const float* __restrict__ x, const float* __restrict__ z,
                                 float* __restrict__ out, long long n, float sqrt_b, float sqrt_a, float c_x0,
                                 float c_xt, float sigma) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const float xv = x[i];
        float x0 = (xv - sqrt_b * eps[i]) / sqrt_a;
        x0 = fminf(fmaxf(x0, -1.f), 1.f);
        float prev = c_x0 * x0 + c_xt * xv;
        if (z) prev = prev + sigma * z[i];
        out[i] = prev;
    }
}
void ddpm_step(const float* eps, const float* x, const float* z, float* out, long long n, float sqrt_b, float sqrt_a,
               float c_x0, float c_xt, float sigma, cudaStream_t s) {
    const int blocks = (int)((n + 255) / 256 < 148 * 8 ? (n + 255) / 256 : 148 * 8);
    ddpm_step_kernel<<<blocks, 256, 0, s>>>(eps, x, z, out, n, sqrt_b, sqrt_a, c_x0, c_xt, sigma);
    SYNT_LAUNCH_CHECK();
}

// =============================================================== time embedding =====
__global__ void __launch_bounds__(256) time_embed_kernel(const float* __restrict__ freqs, const float* __restrict__ w1,
                                                         const float* __restrict__ b1, const float* __restrict__ w2,
                                                         const float* __restrict__ b2, float* __restrict__ emb_silu) {
    __shared__ float e64[64];
    __shared__ float h1[256];
    const int t = blockIdx.x, j = threadIdx.x;
    if (j < 64) {
        const float arg = (float)t * freqs[j & 31];
        e64[j] = j < 32 ? cosf(arg) : sinf(arg);         // flip_sin_to_cos=True -> [cos | sin]
    }
    __syncthreads();
    float a = b1[j];
    for (int i = 0; i < 64; ++i) a = fmaf(w1[j * 64 + i], e64[i], a);
    h1[j] = silu_precise(a);
    __syncthreads();
    float o = b2[j];
    for (int i = 0; i < 256; ++i) o = fmaf(w2[j * 256 + i], h1[i], o);
    emb_silu[(size_t)t * 256 + j] = silu_precise(o);     // every consumer applies SiLU first
}
void time_embed_table(const float* freqs32, const float* w1, const float* b1, const float* w2, const float* b2, int T,
                      float* emb_silu, cudaStream_t s) {
    time_embed_kernel<<<T, 256, 0, s>>>(freqs32, w1, b1, w2, b2, emb_silu);
    SYNT_LAUNCH_CHECK();
}

constexpr int TP_TT = 8;
__global__ void __launch_bounds__(256) time_proj_kernel(const float* __restrict__ emb_silu, const float* __restrict__ w,
                                                        const float* __restrict__ b, int T, int ntot,
                                                        float* __restrict__ table) {
    __shared__ float e[TP_TT][256];
    const int t0 = blockIdx.y * TP_TT;
    for (int i = threadIdx.x; i < TP_TT * 256; i += 256) {
        const int tt = i / 256;
        e[tt][i % 256] = (t0 + tt < T) ? emb_silu[(size_t)(t0 + tt) * 256 + i % 256] : 0.f;
    }
    __syncthreads();
    const int n = blockIdx.x * 256 + threadIdx.x;
    if (n >= ntot) return;
    float acc[TP_TT];
#pragma unroll
    for (int tt = 0; tt < TP_TT; ++tt) acc[tt] = b[n];
    const float4* wr = reinterpret_cast<const float4*>(w + (size_t)n * 256);
    for (int i4 = 0; i4 < 64; ++i4) {
        const float4 wv = __ldg(wr + i4);
#pragma unroll
        for (int tt = 0; tt < TP_TT; ++tt) {
            acc[tt] = fmaf(wv.x, e[tt][4 * i4 + 0], acc[tt]);
            acc[tt] = fmaf(wv.y, e[tt][4 * i4 + 1], acc[tt]);
            acc[tt] = fmaf(wv.z, e[tt][4 * i4 + 2], acc[tt]);
            acc[tt] = fmaf(wv.w, e[tt][4 * i4 + 3], acc[tt]);
        }
    }
#pragma unroll
    for (int tt = 0; tt < TP_TT; ++tt)
        if (t0 + tt < T) table[(size_t)(t0 + tt) * ntot + n] = acc[tt];
}
void time_proj_table(const float* emb_silu, const float* w, const float* b, int T, int ntot, float* table,
                     cudaStream_t s) {
    dim3 grid(ceil_div(ntot, 256), ceil_div(T, TP_TT));
    time_proj_kernel<<<grid, 256, 0, s>>>(emb_silu, w, b, T, ntot, table);
    SYNT_LAUNCH_CHECK();
}

__global__ void select_timestep_kernel(const float* __restrict__ table, int ntot, const float* __restrict__ coef_table,
                                       const int* __restrict__ timesteps, const int* __restrict__ step_ptr, int t_direct,
                                       float* __restrict__ temb_cur, float* __restrict__ coef_cur) {
    pdl_enter();
    const int step = step_ptr ? *step_ptr : 0;
    const int t = step_ptr ? timesteps[step] : t_direct;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < ntot; i += gridDim.x * blockDim.x)
        temb_cur[i] = table[(size_t)t * ntot + i];
    if (coef_table && coef_cur && blockIdx.x == 0 && threadIdx.x < 5)
        coef_cur[threadIdx.x] = coef_table[(size_t)step * 5 + threadIdx.x];
}
void select_timestep(const float* table, int ntot, const float* coef_table, const int* timesteps, const int* step_ptr,
                     int t_direct, float* temb_cur, float* coef_cur, cudaStream_t s) {
    launch_pdl(select_timestep_kernel, dim3(ceil_div(ntot, 256)), dim3(256), 0, s, table, ntot, coef_table, timesteps, step_ptr,
               t_direct, temb_cur, coef_cur);
}
__global__ void advance_step_kernel(int* p) { pdl_enter(); *p += 1; }
__global__ void set_step_kernel(int* p, int v) { *p = v; }
void set_step(int* step_ptr, int value, cudaStream_t s) {
    set_step_kernel<<<1, 1, 0, s>>>(step_ptr, value);
    SYNT_LAUNCH_CHECK();
}
void advance_step(int* step_ptr, cudaStream_t s) {
    launch_pdl(advance_step_kernel, dim3(1), dim3(1), 0, s, step_ptr);
}

// =============================================================== format helpers =====
template <typename T>
__global__ void nhwc_to_nchw_kernel(const T* __restrict__ in, int HW, int C, long long n, float* __restrict__ out) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(i % C);
        const long long p = i / C;                 // b*HW + pix
        const long long b = p / HW, pix = p % HW;
        out[(b * C + c) * HW + pix] = to_f<T>(in[i]);
    }
}
void nhwc_to_nchw_f32(const void* in, int dt, int B, int HW, int C, float* out, cudaStream_t s) {
    const long long n = (long long)B * HW * C;
    const int blocks = (int)((n + 255) / 256 < 148 * 16 ? (n + 255) / 256 : 148 * 16);
    if (dt == DT_F32) nhwc_to_nchw_kernel<float><<<blocks, 256, 0, s>>>((const float*)in, HW, C, n, out);
    else              nhwc_to_nchw_kernel<bf16><<<blocks, 256, 0, s>>>((const bf16*)in, HW, C, n, out);
    SYNT_LAUNCH_CHECK();
}

// mode 0: trunc(clamp((x+1)/2, 0, 1) * 255)      core/generator/image_generator.py:441-447
// mode 1: trunc(clip((x+1)*127.5, 0, 255))       diffusion/diffusion_generator.py:147-148
__global__ void to_uint8_kernel(const float* __restrict__ x, int HW, long long n, int mode, unsigned char* __restrict__ out) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(i % 3);
        const long long p = i / 3;
        const long long b = p / HW, pix = p % HW;
        const float v = x[(b * 3 + c) * HW + pix];
        float r;
        if (mode == 0) { float u = (v + 1.f) / 2.f; u = fminf(fmaxf(u, 0.f), 1.f); r = u * 255.f; }
        else           { r = fminf(fmaxf((v + 1.f) * 127.5f, 0.f), 255.f); }
        out[i] = (unsigned char)r;                 // C-style truncation == numpy astype(uint8)
    }
}
void to_uint8_hwc(const float* x_nchw, int B, int H, int W, int mode, unsigned char* out, cudaStream_t s) {
    const long long n = (long long)B * H * W * 3;
    const int blocks = (int)((n + 255) / 256 < 148 * 8 ? (n + 255) / 256 : 148 * 8);
    to_uint8_kernel<<<blocks, 256, 0, s>>>(x_nchw, H * W, n, mode, out);
    SYNT_LAUNCH_CHECK();
}

}  // namespace synt
