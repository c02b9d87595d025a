// Input-gradient path of the ResNet18 classifier: d log(p_c + 1e-8) / d x for a batch of images, the quantity that
// Integrated Gradients (xai/XAI.py:1039-1085, captum `IntegratedGradients.attribute(..., method='riemann_right')`) and the
// plain gradient attribution (xai/XAI.py:1087-1109) evaluate.  SURVEY.md section 8f row 2.
//
// The data-gradient of every BN-folded convolution is itself a stride-1 convolution (flipped taps, channel roles swapped;
// stride-2 layers through a zero-inserted gradient plane), so the heavy work runs on the SAME implicit-GEMM kernels as the
// forward pass (conv_tc2 / conv_tc on tcgen05, conv_simt in fp32 verification mode) with weight matrices re-arranged once
// on the device (dgrad_weights).  This file holds what is left: the small HBM-bound kernels between the convolutions.
//
//   score_grad        softmax -> s = log(p_c + 1e-8), ds/dlogits                              (one thread per image)
//   avgpool_fc_bwd    dlogits -> gradient at the last block output, ReLU mask fused            [B,7,7,512]
//   relu_mask         g * (act > 0)                                                            (in place)
//   zero_insert       [B,Ho,Wo,C] -> [B,2Ho,2Wo,C] with the values at even positions (+ optional ReLU mask)
//   maxpool_idx / maxpool_bwd   3x3/s2 max-pool that records the winning window element (first maximum, like ATen) and its
//                     adjoint as a gather over the recorded codes, ReLU mask of the stem fused
//   stem_dgrad        7x7/s2 data gradient 64 -> 3 channels on CUDA cores (0.24 GFLOP per image), fp32 output
//   preprocess_bwd    adjoint of clamp((x+1)/2) -> bilinear 128->224 -> normalise, as a gather (deterministic)
//   ig_interpolate / ig_reduce   the Riemann-right path points and the final (x - x') * mean(grad)
#include "kernels.cuh"
#include <cstdlib>

namespace synt {

__global__ void score_grad_kernel(const float* __restrict__ logits, int B, int nc, int target, float* __restrict__ score,
                                  float* __restrict__ dlogits) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    const float* l = logits + (size_t)b * nc;
    float m = l[0];
    for (int k = 1; k < nc; ++k) m = fmaxf(m, l[k]);
    float z = 0.f;
    for (int k = 0; k < nc; ++k) z += expf(l[k] - m);
    const float pc = expf(l[target] - m) / z;
    if (score) score[b] = logf(pc + 1e-8f);
    const float f = pc / (pc + 1e-8f);                       // d log(p_c + eps) / d p_c * p_c
    for (int k = 0; k < nc; ++k) {
        const float pk = expf(l[k] - m) / z;
        dlogits[(size_t)b * nc + k] = f * ((k == target ? 1.f : 0.f) - pk);
    }
}
void score_grad(const float* logits, int B, int nc, int target, float* score, float* dlogits, cudaStream_t s) {
    SYNT_CHECK(target >= 0 && target < nc, "score_grad: target class out of range");
    score_grad_kernel<<<(B + 127) / 128, 128, 0, s>>>(logits, B, nc, target, score, dlogits);
    SYNT_LAUNCH_CHECK();
}

template <typename T>
__global__ void __launch_bounds__(512) avgpool_fc_bwd_kernel(const float* __restrict__ dlogits, const float* __restrict__ w,
                                                             const T* __restrict__ feat, int HW, int C, int nc,
                                                             T* __restrict__ out) {
    const int b = blockIdx.x, c = threadIdx.x;
    if (c >= C) return;
    float v = 0.f;
    for (int k = 0; k < nc; ++k) v = fmaf(dlogits[(size_t)b * nc + k], w[(size_t)k * C + c], v);
    v /= (float)HW;
    for (int p = 0; p < HW; ++p) {
        const size_t i = ((size_t)b * HW + p) * C + c;
        out[i] = from_f<T>(to_f<T>(feat[i]) > 0.f ? v : 0.f);
    }
}
void avgpool_fc_bwd(const float* dlogits, const float* w, const void* feat, int dt, int B, int HW, int C, int nc, void* out,
                    cudaStream_t s) {
    SYNT_CHECK(C <= 512, "avgpool_fc_bwd: C <= 512");
    if (dt == DT_F32) avgpool_fc_bwd_kernel<float><<<B, 512, 0, s>>>(dlogits, w, (const float*)feat, HW, C, nc, (float*)out);
    else              avgpool_fc_bwd_kernel<bf16><<<B, 512, 0, s>>>(dlogits, w, (const bf16*)feat, HW, C, nc, (bf16*)out);
    SYNT_LAUNCH_CHECK();
}

static inline int ew_blocks(long long n) { return (int)((n + 255) / 256 < 148 * 16 ? (n + 255) / 256 : 148 * 16); }

// out may alias g
template <typename T>
__global__ void relu_mask_kernel(const T* g, const T* __restrict__ act, long long nvec, T* out) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < nvec; i += (long long)gridDim.x * blockDim.x) {
        float a[8], v[8];
        load8<T>(act + i * 8, a);
        load8<T>(g + i * 8, v);
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = a[j] > 0.f ? v[j] : 0.f;
        store8<T>(out + i * 8, v);
    }
}
void relu_mask(const void* g, const void* act, int dt, long long n, void* out, cudaStream_t s) {
    SYNT_CHECK(n % 8 == 0, "relu_mask: element count must be a multiple of 8");
    const long long nv = n / 8;
    if (dt == DT_F32) relu_mask_kernel<float><<<ew_blocks(nv), 256, 0, s>>>((const float*)g, (const float*)act, nv, (float*)out);
    else              relu_mask_kernel<bf16><<<ew_blocks(nv), 256, 0, s>>>((const bf16*)g, (const bf16*)act, nv, (bf16*)out);
    SYNT_LAUNCH_CHECK();
}

template <typename T>
__global__ void zero_insert_kernel(const T* __restrict__ g, const T* __restrict__ act /* nullable */, int Ho, int Wo, int C,
                                   long long nvec_out, T* __restrict__ out) {
    const int nvec = C >> 3, W2 = 2 * Wo, H2 = 2 * Ho;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < nvec_out; i += (long long)gridDim.x * blockDim.x) {
        const int v = (int)(i % nvec);
        long long p = i / nvec;
        const int x = (int)(p % W2); p /= W2;
        const int y = (int)(p % H2);
        const long long b = p / H2;
        float o[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = 0.f;
        if (!((x | y) & 1)) {
            const long long src = ((b * Ho + (y >> 1)) * Wo + (x >> 1)) * C + v * 8;
            load8<T>(g + src, o);
            if (act) {
                float a[8];
                load8<T>(act + src, a);
#pragma unroll
                for (int j = 0; j < 8; ++j) o[j] = a[j] > 0.f ? o[j] : 0.f;
            }
        }
        store8<T>(out + i * 8, o);
    }
}
void zero_insert2x(const void* g, const void* act, int dt, int B, int Ho, int Wo, int C, void* out, cudaStream_t s) {
    SYNT_CHECK(C % 8 == 0, "zero_insert2x: C must be a multiple of 8");
    const long long nv = (long long)B * 4 * Ho * Wo * (C / 8);
    if (dt == DT_F32) zero_insert_kernel<float><<<ew_blocks(nv), 256, 0, s>>>((const float*)g, (const float*)act, Ho, Wo, C, nv, (float*)out);
    else              zero_insert_kernel<bf16><<<ew_blocks(nv), 256, 0, s>>>((const bf16*)g, (const bf16*)act, Ho, Wo, C, nv, (bf16*)out);
    SYNT_LAUNCH_CHECK();
}

// 3x3 / stride 2 / pad 1 max-pool that also records WHICH window element won: code = dy*3 + dx of the FIRST maximum in
// scan order (ATen's max_pool2d keeps the first maximum: `val > maxval`), one byte per output element.
template <typename T>
__global__ void maxpool_idx_kernel(const T* __restrict__ in, int H, int W, int C, int Ho, int Wo, long long nvec_total,
                                   T* __restrict__ out, unsigned char* __restrict__ idx) {
    const int nvec = C >> 3;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < nvec_total; i += (long long)gridDim.x * blockDim.x) {
        const int v = (int)(i % nvec);
        long long p = i / nvec;
        const int ox = (int)(p % Wo); p /= Wo;
        const int oy = (int)(p % Ho);
        const long long b = p / Ho;
        float m[8];
        unsigned int code[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) { m[j] = -INFINITY; code[j] = 0; }
#pragma unroll
        for (int dy = 0; dy < 3; ++dy) {
            const int iy = oy * 2 - 1 + dy;
            if (iy < 0 || iy >= H) continue;
#pragma unroll
            for (int dx = 0; dx < 3; ++dx) {
                const int ix = ox * 2 - 1 + dx;
                if (ix < 0 || ix >= W) continue;
                float x[8];
                load8<T>(in + ((b * H + iy) * W + ix) * C + v * 8, x);
#pragma unroll
                for (int j = 0; j < 8; ++j)
                    if (x[j] > m[j]) { m[j] = x[j]; code[j] = dy * 3 + dx; }
            }
        }
        store8<T>(out + i * 8, m);
        uint2 pk;
        pk.x = code[0] | (code[1] << 8) | (code[2] << 16) | (code[3] << 24);
        pk.y = code[4] | (code[5] << 8) | (code[6] << 16) | (code[7] << 24);
        *reinterpret_cast<uint2*>(idx + i * 8) = pk;
    }
}
void maxpool3x3s2_idx(const void* in, int dt, int B, int H, int W, int C, void* out, unsigned char* idx, cudaStream_t s) {
    const int Ho = (H + 2 - 3) / 2 + 1, Wo = (W + 2 - 3) / 2 + 1;
    const long long nv = (long long)B * Ho * Wo * (C / 8);
    if (dt == DT_F32) maxpool_idx_kernel<float><<<ew_blocks(nv), 256, 0, s>>>((const float*)in, H, W, C, Ho, Wo, nv, (float*)out, idx);
    else              maxpool_idx_kernel<bf16><<<ew_blocks(nv), 256, 0, s>>>((const bf16*)in, H, W, C, Ho, Wo, nv, (bf16*)out, idx);
    SYNT_LAUNCH_CHECK();
}

// Adjoint of the max-pool as a gather: input pixel (iy, ix) collects dpool of each of the (at most four) windows whose
// recorded winner is this pixel.  The ReLU mask of the stem output is applied on the way out (a zero activation gets no
// gradient either way).  Per thread: one 16-byte activation load, <= 4 x (8-byte code + 16-byte gradient) loads.
template <typename T>
__global__ void maxpool_bwd_kernel(const T* __restrict__ dpool, const unsigned char* __restrict__ idx, const T* __restrict__ act,
                                   int H, int W, int C, int Ho, int Wo, long long nvec_total, T* __restrict__ dact) {
    const int nvec = C >> 3;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < nvec_total; i += (long long)gridDim.x * blockDim.x) {
        const int v = (int)(i % nvec);
        long long p = i / nvec;
        const int ix = (int)(p % W); p /= W;
        const int iy = (int)(p % H);
        const long long b = p / H;
        float self[8], acc[8];
        load8<T>(act + i * 8, self);
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] = 0.f;
        const int oy_lo = iy >> 1, oy_hi = (iy + 1) >> 1;           // windows 2*oy-1 .. 2*oy+1 that contain iy
        const int ox_lo = ix >> 1, ox_hi = (ix + 1) >> 1;
        for (int oy = oy_lo; oy <= oy_hi; ++oy) {
            if (oy >= Ho) continue;
            for (int ox = ox_lo; ox <= ox_hi; ++ox) {
                if (ox >= Wo) continue;
                const unsigned int mine = (unsigned int)((iy - (2 * oy - 1)) * 3 + (ix - (2 * ox - 1)));
                const long long o = ((b * Ho + oy) * Wo + ox) * C + v * 8;
                const uint2 pk = *reinterpret_cast<const uint2*>(idx + o);
                float d[8];
                load8<T>(dpool + o, d);
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const unsigned int code = ((j < 4 ? pk.x : pk.y) >> (8 * (j & 3))) & 0xffu;
                    acc[j] += code == mine ? d[j] : 0.f;
                }
            }
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] = self[j] > 0.f ? acc[j] : 0.f;
        store8<T>(dact + i * 8, acc);
    }
}
void maxpool3x3s2_bwd(const void* dpool, const unsigned char* idx, const void* act, int dt, int B, int H, int W, int C, void* dact,
                      cudaStream_t s) {
    const int Ho = (H + 2 - 3) / 2 + 1, Wo = (W + 2 - 3) / 2 + 1;
    const long long nv = (long long)B * H * W * (C / 8);
    if (dt == DT_F32) maxpool_bwd_kernel<float><<<ew_blocks(nv), 256, 0, s>>>((const float*)dpool, idx, (const float*)act, H, W, C, Ho, Wo, nv, (float*)dact);
    else              maxpool_bwd_kernel<bf16><<<ew_blocks(nv), 256, 0, s>>>((const bf16*)dpool, idx, (const bf16*)act, H, W, C, Ho, Wo, nv, (bf16*)dact);
    SYNT_LAUNCH_CHECK();
}

// Data gradient of the 7x7 / stride 2 / pad 3 stem (3 -> 64 channels, BN folded): dpre[iy, ix, c] =
// sum over (ky, kx, co) with 2*oy = iy + 3 - ky, 2*ox = ix + 3 - kx of g[oy, ox, co] * w[co][(ky*7 + kx)*3 + c].
// One CTA works on ONE parity class (iy % 2, ix % 2): every thread then uses the same 3x3 .. 4x4 subset of taps, so the
// weights are a shared-memory broadcast (one LDS.128 per output channel and tap).  A thread owns FOUR horizontally adjacent
// pixels of the class, so each broadcast weight feeds 12 FMAs (0.24 GFLOP per image on CUDA cores; with one pixel per
// thread the kernel was bound by the LDS issue rate: 1.55 ms -> see profiles/r01f).  g is already masked by the stem's ReLU.
constexpr int SD_PX = 4, SD_GROUPS = 112 / SD_PX;                       // 28 pixel groups per row of a parity class
template <typename T>
__global__ void __launch_bounds__(256) stem_dgrad_kernel(const T* __restrict__ g, const float* __restrict__ w /* [64][147] */,
                                                         float* __restrict__ dpre /* [B,224,224,3] */) {
    __shared__ float4 ws[16][64];
    const int par_y = blockIdx.y >> 1, par_x = blockIdx.y & 1;
    const long long b = blockIdx.z;
    // taps of this parity class: ky = ky0 + 2a with (iy + 3 - ky) even  ->  ky0 = (par_y + 1) & 1
    const int ky0 = (par_y + 1) & 1, kx0 = (par_x + 1) & 1;
    const int nky = ky0 ? 3 : 4, nkx = kx0 ? 3 : 4;
    for (int i = threadIdx.x; i < nky * nkx * 64; i += 256) {
        const int co = i % 64, t = i / 64;
        const int ky = ky0 + 2 * (t / nkx), kx = kx0 + 2 * (t % nkx);
        const float* wp = w + co * 147 + (ky * 7 + kx) * 3;
        ws[t][co] = make_float4(wp[0], wp[1], wp[2], 0.f);
    }
    __syncthreads();
    const int q = blockIdx.x * 256 + threadIdx.x;
    if (q >= 112 * SD_GROUPS) return;
    const int qy = q / SD_GROUPS, qx0 = (q % SD_GROUPS) * SD_PX;
    const int iy = 2 * qy + par_y;
    float acc[SD_PX][3];
#pragma unroll
    for (int j = 0; j < SD_PX; ++j) acc[j][0] = acc[j][1] = acc[j][2] = 0.f;
    for (int a = 0; a < nky; ++a) {
        const int ny = iy + 3 - (ky0 + 2 * a);                        // even by construction
        if (ny < 0 || (ny >> 1) >= 112) continue;
        const int oy = ny >> 1;
        for (int e = 0; e < nkx; ++e) {
            // ox of pixel j = (2*(qx0 + j) + par_x + 3 - kx) / 2 = qx0 + j + sh, sh = (par_x + 3 - kx) / 2 (exact, may be -1)
            const int sh = (par_x + 3 - (kx0 + 2 * e)) >> 1;
            const int ox0 = qx0 + sh;
            const T* gp = g + ((b * 112 + oy) * 112 + ox0) * 64;
            const float4* wt = ws[a * nkx + e];
#pragma unroll
            for (int v = 0; v < 8; ++v) {
                float x[SD_PX][8];
#pragma unroll
                for (int j = 0; j < SD_PX; ++j) {
                    if (ox0 + j >= 0 && ox0 + j < 112) load8<T>(gp + j * 64 + v * 8, x[j]);
                    else {
#pragma unroll
                        for (int k = 0; k < 8; ++k) x[j][k] = 0.f;
                    }
                }
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    const float4 wv = wt[v * 8 + k];
#pragma unroll
                    for (int j = 0; j < SD_PX; ++j) {
                        acc[j][0] = fmaf(x[j][k], wv.x, acc[j][0]);
                        acc[j][1] = fmaf(x[j][k], wv.y, acc[j][1]);
                        acc[j][2] = fmaf(x[j][k], wv.z, acc[j][2]);
                    }
                }
            }
        }
    }
#pragma unroll
    for (int j = 0; j < SD_PX; ++j) {
        float* o = dpre + ((b * 224 + iy) * 224 + 2 * (qx0 + j) + par_x) * 3;
        o[0] = acc[j][0]; o[1] = acc[j][1]; o[2] = acc[j][2];
    }
}
// bf16 production variant: the kernel above re-reads every g row once per tap and parity class (49x in total, ~10 GB of
// L2 -> L1 traffic per 64 images: 1.36 ms, L2-bandwidth bound).  Here one CTA owns a 32x32 block of dpre (all four parity
// classes, 16x16 pixels each), stages the 19x19x64 g tile it needs ONCE in shared memory (zero-filled outside the plane, so
// the tap loops need no bounds checks) next to the 49 x 64 weight triples, and every tap reads from there.
// Pixel pitch 144 B and row pitch 2752 B (= 4 mod 8 sixteen-byte units) keep the LDS.128 of 8 adjacent lanes (4 pixels x 2
// rows) on distinct bank groups; a thread owns 4 pixels of one class at column stride 4; warps are class-pure, so the weight
// reads are broadcasts.  102 KB of shared memory -> two CTAs per SM.
constexpr int SDT_PITCH = 144, SDT_ROW = 19 * SDT_PITCH + 16, SDT_TILE_BYTES = 19 * SDT_ROW;      // 2752, 52288
constexpr int SDT_W_BYTES = 49 * 64 * 16, SDT_SMEM = SDT_TILE_BYTES + SDT_W_BYTES;                   // 50176, 102464
__device__ __forceinline__ int sdt_class_base(int cls) { return cls == 0 ? 0 : (cls == 1 ? 9 : (cls == 2 ? 21 : 33)); }
__global__ void __launch_bounds__(256, 2) stem_dgrad_tiled_kernel(const bf16* __restrict__ g, const float* __restrict__ w,
                                                                  float* __restrict__ dpre) {
    extern __shared__ __align__(16) unsigned char sdt_smem[];
    unsigned char* tile = sdt_smem;
    float4* ws = reinterpret_cast<float4*>(sdt_smem + SDT_TILE_BYTES);
    const int tyi = blockIdx.x / 7, txi = blockIdx.x % 7;
    const long long b = blockIdx.y;
    const int oy_base = tyi * 16 - 1, ox_base = txi * 16 - 1;
    // weights: class cls = 2*(iy%2) + (ix%2) uses taps ky = ky0 + 2a, kx = kx0 + 2e, stored at [base(cls) + a*nkx + e][co]
    for (int i = threadIdx.x; i < 49 * 64; i += 256) {
        const int co = i & 63, t = i >> 6;
        const int cls = t < 9 ? 0 : (t < 21 ? 1 : (t < 33 ? 2 : 3));
        const int tl = t - sdt_class_base(cls);
        const int ky0 = ((cls >> 1) + 1) & 1, kx0 = ((cls & 1) + 1) & 1, nkx = kx0 ? 3 : 4;
        const int ky = ky0 + 2 * (tl / nkx), kx = kx0 + 2 * (tl % nkx);
        const float* wp = w + co * 147 + (ky * 7 + kx) * 3;
        ws[i] = make_float4(wp[0], wp[1], wp[2], 0.f);
    }
    for (int i = threadIdx.x; i < 19 * 19 * 8; i += 256) {
        const int ch = i & 7, p = i >> 3;
        const int ly = p / 19, lx = p - ly * 19;
        const int oy = oy_base + ly, ox = ox_base + lx;
        uint4 val = make_uint4(0u, 0u, 0u, 0u);
        if (oy >= 0 && oy < 112 && ox >= 0 && ox < 112)
            val = *reinterpret_cast<const uint4*>(g + ((b * 112 + oy) * 112 + ox) * 64 + ch * 8);
        *reinterpret_cast<uint4*>(tile + ly * SDT_ROW + lx * SDT_PITCH + ch * 16) = val;
    }
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int cls = warp >> 1, par_y = cls >> 1, par_x = cls & 1;
    const int t = (warp & 1) * 32 + lane;                           // 64 threads per class: 16 rows x 4 column phases
    const int qy = t >> 2, qxg = t & 3;                             // the thread's pixels: class row qy, columns qxg + 4j
    const int ky0 = (par_y + 1) & 1, kx0 = (par_x + 1) & 1;
    const int nky = ky0 ? 3 : 4, nkx = kx0 ? 3 : 4;
    const float4* wcls = ws + sdt_class_base(cls) * 64;
    float acc[4][3];
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[j][0] = acc[j][1] = acc[j][2] = 0.f;
    for (int a = 0; a < nky; ++a) {
        const int ly = qy + ((par_y + 3 - (ky0 + 2 * a)) >> 1) + 1;                 // oy - oy_base, 0..18
        for (int e = 0; e < nkx; ++e) {
            const int lx = qxg + ((par_x + 3 - (kx0 + 2 * e)) >> 1) + 1;             // column of pixel j = 0
            const unsigned char* src = tile + ly * SDT_ROW + lx * SDT_PITCH;
            const float4* wt = wcls + (a * nkx + e) * 64;
#pragma unroll
            for (int v = 0; v < 8; ++v) {
                float x[4][8];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const uint4 r = *reinterpret_cast<const uint4*>(src + j * 4 * SDT_PITCH + v * 16);
                    const unsigned int u[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        x[j][2 * k] = __uint_as_float(u[k] << 16);
                        x[j][2 * k + 1] = __uint_as_float(u[k] & 0xffff0000u);
                    }
                }
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    const float4 wv = wt[v * 8 + k];
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        acc[j][0] = fmaf(x[j][k], wv.x, acc[j][0]);
                        acc[j][1] = fmaf(x[j][k], wv.y, acc[j][1]);
                        acc[j][2] = fmaf(x[j][k], wv.z, acc[j][2]);
                    }
                }
            }
        }
    }
    const int iy = tyi * 32 + 2 * qy + par_y;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int ix = txi * 32 + 2 * (qxg + 4 * j) + par_x;
        float* o = dpre + ((b * 224 + iy) * 224 + ix) * 3;
        o[0] = acc[j][0]; o[1] = acc[j][1]; o[2] = acc[j][2];
    }
}

void stem_dgrad(const void* g, int dt, int B, const float* w, float* dpre, cudaStream_t s) {
    static const bool tiled = [] { const char* e = getenv("SYNT_STEM_DGRAD_TILED"); return !(e && e[0] == '0'); }();
    if (dt == DT_BF16 && tiled) {
        ensure_dynamic_smem((const void*)(stem_dgrad_tiled_kernel), SDT_SMEM);
        stem_dgrad_tiled_kernel<<<dim3(49, B), 256, SDT_SMEM, s>>>((const bf16*)g, w, dpre);
        SYNT_LAUNCH_CHECK();
        return;
    }
    dim3 grid((112 * SD_GROUPS + 255) / 256, 4, B);
    if (dt == DT_F32) stem_dgrad_kernel<float><<<grid, 256, 0, s>>>((const float*)g, w, dpre);
    else              stem_dgrad_kernel<bf16><<<grid, 256, 0, s>>>((const bf16*)g, w, dpre);
    SYNT_LAUNCH_CHECK();
}

// Adjoint of preprocess_pixel (resnet.cu): dx[b,c,y,x] = 0.5 * [0 <= (x+1)/2 <= 1] / std[c] *
// sum over (oy, ox) of wy(oy, y) * wx(ox, x) * dpre[b, oy, ox, c], with the SAME source-index arithmetic as the forward
// kernel (fy = max((oy + 0.5) * sy - 0.5, 0), y0 = floor, y1 = min(y0 + 1, Hin - 1)).  Gather -> deterministic.
__global__ void preprocess_bwd_kernel(const float* __restrict__ dpre, const float* __restrict__ x, int Hin, int Win, int Hout,
                                      int Wout, long long npix, float* __restrict__ dx) {
    const float stdv[3] = {0.229f, 0.224f, 0.225f};
    const float sy = (float)Hin / (float)Hout, sx = (float)Win / (float)Wout;
    constexpr int NC = 8;                                     // candidates per axis (scale 1.75 -> at most 5 contribute)
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < npix; i += (long long)gridDim.x * blockDim.x) {
        const int xx = (int)(i % Win), yy = (int)((i / Win) % Hin);
        const long long b = i / ((long long)Win * Hin);
        auto weights = [](int pos, int n_in, int n_out, float sc, int& lo, float (&wgt)[NC]) {
            lo = (int)floorf(((float)pos - 0.5f) / sc - 0.5f) - 1;
            if (lo < 0) lo = 0;
#pragma unroll
            for (int k = 0; k < NC; ++k) {
                const int o = lo + k;
                float wv = 0.f;
                if (o < n_out) {
                    float f = ((float)o + 0.5f) * sc - 0.5f; if (f < 0.f) f = 0.f;
                    const int p0 = (int)f, p1 = p0 + (p0 < n_in - 1 ? 1 : 0);
                    const float l = f - (float)p0;
                    if (pos == p0) wv += 1.f - l;
                    if (pos == p1) wv += l;
                }
                wgt[k] = wv;
            }
        };
        int ylo, xlo;
        float wy[NC], wx[NC];
        weights(yy, Hin, Hout, sy, ylo, wy);
        weights(xx, Win, Wout, sx, xlo, wx);
        float acc[3] = {0.f, 0.f, 0.f};
#pragma unroll
        for (int a = 0; a < NC; ++a) {
            if (wy[a] == 0.f) continue;
#pragma unroll
            for (int e = 0; e < NC; ++e) {
                if (wx[e] == 0.f) continue;
                const float wgt = wy[a] * wx[e];
                const float* p = dpre + ((b * Hout + ylo + a) * Wout + xlo + e) * 3;
                acc[0] = fmaf(wgt, p[0], acc[0]); acc[1] = fmaf(wgt, p[1], acc[1]); acc[2] = fmaf(wgt, p[2], acc[2]);
            }
        }
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const long long idx = ((b * 3 + c) * Hin + yy) * Win + xx;
            const float v = (x[idx] + 1.0f) / 2.0f;
            dx[idx] = (v >= 0.f && v <= 1.f) ? 0.5f * acc[c] / stdv[c] : 0.f;   // clamp passes the gradient on [min, max]
        }
    }
}
void classifier_preprocess_bwd(const float* dpre, const float* x, int B, int Hin, int Win, int Hout, int Wout, float* dx,
                               cudaStream_t s) {
    const long long npix = (long long)B * Hin * Win;
    preprocess_bwd_kernel<<<ew_blocks(npix), 256, 0, s>>>(dpre, x, Hin, Win, Hout, Wout, npix, dx);
    SYNT_LAUNCH_CHECK();
}

// dst[n][dst_off + t*cout_f + co] = src[co][src_off + (taps-1-t)*cin_f + n]: the K-major weight matrix of the
// data-gradient convolution (flipped taps, channel roles swapped) from the forward matrix.
template <typename T>
__global__ void dgrad_weight_kernel(const T* __restrict__ src, int src_ld, int src_off, int cin_f, int cout_f, int taps,
                                    T* __restrict__ dst, int dst_ld, int dst_off, long long n) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const int co = (int)(i % cout_f);
        const int t = (int)((i / cout_f) % taps);
        const int ci = (int)(i / ((long long)cout_f * taps));
        dst[(size_t)ci * dst_ld + dst_off + t * cout_f + co] = src[(size_t)co * src_ld + src_off + (taps - 1 - t) * cin_f + ci];
    }
}
void dgrad_weights(const void* src, int src_ld, int src_off, int cin_f, int cout_f, int taps, int bf, void* dst, int dst_ld,
                   int dst_off, cudaStream_t s) {
    const long long n = (long long)cin_f * taps * cout_f;
    if (bf) dgrad_weight_kernel<unsigned short><<<ew_blocks(n), 256, 0, s>>>((const unsigned short*)src, src_ld, src_off, cin_f, cout_f, taps,
                                                                           (unsigned short*)dst, dst_ld, dst_off, n);
    else    dgrad_weight_kernel<float><<<ew_blocks(n), 256, 0, s>>>((const float*)src, src_ld, src_off, cin_f, cout_f, taps,
                                                                   (float*)dst, dst_ld, dst_off, n);
    SYNT_LAUNCH_CHECK();
}

// Integrated Gradients plumbing (captum riemann_right: alpha_k = k/n, k = 1..n, step 1/n)
__global__ void ig_interpolate_kernel(const float* __restrict__ x, const float* __restrict__ base, int n_steps, long long per,
                                      float* __restrict__ out) {
    const long long n = per * n_steps;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const long long e = i % per;
        const int k = (int)(i / per);
        const float alpha = (float)((double)(k + 1) / (double)n_steps);
        out[i] = base[e] + alpha * (x[e] - base[e]);
    }
}
void ig_interpolate(const float* x, const float* base, int n_steps, long long per, float* out, cudaStream_t s) {
    ig_interpolate_kernel<<<ew_blocks(per * n_steps), 256, 0, s>>>(x, base, n_steps, per, out);
    SYNT_LAUNCH_CHECK();
}
__global__ void ig_reduce_kernel(const float* __restrict__ grads, const float* __restrict__ x, const float* __restrict__ base,
                                 int n_steps, long long per, float* __restrict__ out) {
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < per; e += (long long)gridDim.x * blockDim.x) {
        const float step = (float)(1.0 / (double)n_steps);
        float acc = 0.f;
        for (int k = 0; k < n_steps; ++k) acc += grads[(long long)k * per + e] * step;
        out[e] = acc * (x[e] - base[e]);
    }
}
void ig_reduce(const float* grads, const float* x, const float* base, int n_steps, long long per, float* out, cudaStream_t s) {
    ig_reduce_kernel<<<ew_blocks(per), 256, 0, s>>>(grads, x, base, n_steps, per, out);
    SYNT_LAUNCH_CHECK();
}

}  // namespace synt
