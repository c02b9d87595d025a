// ResNet18 classifier path behind the C ABI: the logit evaluations that Time-SHAP,
// patch-SHAP and the CSI interventions issue (reference: MelanomaClassifierAdaptive,
// xai/XAI.py:357-471; call sites xai/XAI.py:1143-1170, :1201-1211, :1622-1671).
//
//   preprocess   clamp((x+1)/2,0,1) -> bilinear 128->224 (align_corners=False; antialias is a
//                no-op when upsampling) -> ImageNet normalise          xai/XAI.py:399-431
//   backbone     torchvision resnet18, eval-mode BatchNorm folded into conv weight/bias,
//                ReLU / residual / 1x1-stride-2 downsample fused into the conv epilogue
//   head         global average pool + Linear(512, num_classes)
#include "kernels.cuh"
#include "preprocess.cuh"
#include "pool.h"
#include "../../include/synt_isic.h"
#include <cmath>
#include <cstring>
#include <memory>
#include <string>
#include <vector>

namespace synt {

extern thread_local std::string g_last_error;

template <typename T>
__global__ void preprocess_kernel(const float* __restrict__ x, int Hin, int Win, int Hout, int Wout, long long npix,
                                  T* __restrict__ out) {
    const float sy = (float)Hin / (float)Hout, sx = (float)Win / (float)Wout;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < npix; i += (long long)gridDim.x * blockDim.x) {
        const int ox = (int)(i % Wout), oy = (int)((i / Wout) % Hout);
        const long long b = i / ((long long)Wout * Hout);
        float v3[3];
        preprocess_pixel(x + b * 3 * (long long)Hin * Win, Hin, Win, sy, sx, oy, ox, v3);
#pragma unroll
        for (int c = 0; c < 3; ++c) out[i * 3 + c] = from_f<T>(v3[c]);
    }
}
void classifier_preprocess(const float* x, int B, int Hin, int Win, int Hout, int Wout, void* out, int Cpad, int dt,
                           cudaStream_t s) {
    SYNT_CHECK(Cpad == 3, "classifier_preprocess: only C=3 output");
    const long long npix = (long long)B * Hout * Wout;
    const int blocks = (int)((npix + 255) / 256 < 148 * 16 ? (npix + 255) / 256 : 148 * 16);
    if (dt == DT_F32) preprocess_kernel<float><<<blocks, 256, 0, s>>>(x, Hin, Win, Hout, Wout, npix, (float*)out);
    else              preprocess_kernel<bf16><<<blocks, 256, 0, s>>>(x, Hin, Win, Hout, Wout, npix, (bf16*)out);
    SYNT_LAUNCH_CHECK();
}

// Stem im2col for the tensor-core path: pre [B,224,224,3] bf16 -> A [B,112,112,192] bf16 with
// k = ky*24 + kx*3 + c (7x7 window, stride 2, pad 3; kx = 7 and k >= 168 are zero padding), so every
// window row is one 48-byte, 16-byte-aligned segment copied from 21 contiguous input elements.
// The 7x7/s2 stem then runs as a 1x1 convolution with Cin = 192 on the tcgen05 kernel.
__global__ void __launch_bounds__(256) stem_im2col_kernel(const bf16* __restrict__ pre, long long nseg, bf16* __restrict__ out) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < nseg; i += (long long)gridDim.x * blockDim.x) {
        const int ky = (int)(i & 7);                       // segment 7 = zero tail
        long long p = i >> 3;
        const int ox = (int)(p % 112), oy = (int)((p / 112) % 112);
        const long long b = p / (112 * 112);
        __align__(16) unsigned short v[24];
#pragma unroll
        for (int j = 0; j < 24; ++j) v[j] = 0;
        const int iy = 2 * oy - 3 + ky;
        if (ky < 7 && iy >= 0 && iy < 224) {
            const unsigned short* row = reinterpret_cast<const unsigned short*>(pre) + ((b * 224 + iy) * 224) * 3;
            const int ix0 = 2 * ox - 3;
#pragma unroll
            for (int kx = 0; kx < 7; ++kx) {
                const int ix = ix0 + kx;
                if (ix >= 0 && ix < 224) {
#pragma unroll
                    for (int c = 0; c < 3; ++c) v[kx * 3 + c] = row[ix * 3 + c];
                }
            }
        }
        uint4* dst = reinterpret_cast<uint4*>(out + p * 192 + ky * 24);
        const uint4* src = reinterpret_cast<const uint4*>(v);
        dst[0] = src[0]; dst[1] = src[1]; dst[2] = src[2];
    }
}
void stem_im2col(const void* pre, int B, void* out, cudaStream_t s) {
    const long long nseg = (long long)B * 112 * 112 * 8;
    const int blocks = (int)((nseg + 255) / 256 < 148 * 32 ? (nseg + 255) / 256 : 148 * 32);
    stem_im2col_kernel<<<blocks, 256, 0, s>>>((const bf16*)pre, nseg, (bf16*)out);
    SYNT_LAUNCH_CHECK();
}

// ---- fused stem: conv 7x7 / stride 2 / pad 3 (3 -> 64, BN folded) + ReLU on the legacy tensor-core path, no im2col tensor.
// One CTA = 16x16 output pixels of one image.  The 38x38x3 input patch is staged in shared memory with the pixels of a row
// contiguous (pitch 120 bf16): for k = ky*24 + kx*3 + c the A operand of output pixel (oy, ox) is the contiguous run
// patch[2*oy + ky][6*ox + (kx*3 + c)], so every mma.sync A register is one aligned 32-bit shared-memory load (the 3 slots
// per ky beyond the 21 real ones read the neighbouring pixel against zero weights).  K = 168 -> 11 steps of 16; the weights
// arrive pre-arranged as per-lane B fragments [11][8 n-tiles][32 lanes].  Each warp owns two output rows (two m16 tiles).
constexpr int ST_TILE = 16, ST_ROWS = 2 * ST_TILE + 6, ST_PITCH = 120, ST_STEPS = 11;
constexpr int ST_PATCH_BYTES = ST_ROWS * ST_PITCH * 2;                   // 9120
constexpr int ST_FRAG_BYTES = ST_STEPS * 8 * 32 * 8;                     // 22528
constexpr int ST_OUT_PITCH = 144;                                        // bytes per staged output pixel (64 bf16 + pad)
constexpr int ST_SMEM = ST_PATCH_BYTES + ST_FRAG_BYTES + ST_TILE * ST_TILE * ST_OUT_PITCH;

__global__ void __launch_bounds__(256) stem_mma_kernel(const bf16* __restrict__ pre, const uint2* __restrict__ bfrag,
                                                       const float* __restrict__ bias, bf16* __restrict__ out) {
    extern __shared__ __align__(16) uint8_t st_smem[];
    unsigned short* patch = reinterpret_cast<unsigned short*>(st_smem);
    uint2* frag_s = reinterpret_cast<uint2*>(st_smem + ST_PATCH_BYTES);
    uint8_t* stage = st_smem + ST_PATCH_BYTES + ST_FRAG_BYTES;
    const int b = blockIdx.z, oy0 = blockIdx.y * ST_TILE, ox0 = blockIdx.x * ST_TILE;
    for (int i = threadIdx.x; i < ST_STEPS * 8 * 32; i += 256) frag_s[i] = __ldg(bfrag + i);
    {
        const unsigned short* src = reinterpret_cast<const unsigned short*>(pre) + (size_t)b * 224 * 224 * 3;
        const int iy0 = 2 * oy0 - 3, ix0 = 2 * ox0 - 3;
        for (int i = threadIdx.x; i < ST_ROWS * ST_PITCH; i += 256) {
            const int r = i / ST_PITCH, e = i - r * ST_PITCH;            // e = pixel*3 + channel within the patch row
            const int iy = iy0 + r, ix = ix0 + e / 3;
            patch[i] = (iy >= 0 && iy < 224 && ix >= 0 && ix < 224) ? __ldg(src + ((size_t)iy * 224 + ix) * 3 + e % 3) : (unsigned short)0;
        }
    }
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    float acc[2][8][4];
#pragma unroll
    for (int m = 0; m < 2; ++m)
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
            const float b0 = __ldg(bias + nt * 8 + t * 2), b1 = __ldg(bias + nt * 8 + t * 2 + 1);
            acc[m][nt][0] = b0; acc[m][nt][1] = b1; acc[m][nt][2] = b0; acc[m][nt][3] = b1;
        }
#pragma unroll
    for (int s = 0; s < ST_STEPS; ++s) {
        // k = 16 s + 2t (+8): window row ky and slot kk within it are compile-time per half of the step
        constexpr int dummy = 0; (void)dummy;
        const int k_lo = 16 * s, k_hi = 16 * s + 8;
        const int off_lo = (k_lo / 24) * ST_PITCH + (k_lo % 24) + 2 * t;
        const int off_hi = (k_hi / 24) * ST_PITCH + (k_hi % 24) + 2 * t;
        uint32_t a[2][4];
#pragma unroll
        for (int m = 0; m < 2; ++m) {
            const unsigned short* row = patch + (2 * (warp * 2 + m)) * ST_PITCH;    // output row warp*2 + m
            a[m][0] = *reinterpret_cast<const uint32_t*>(row + 6 * g + off_lo);
            a[m][1] = *reinterpret_cast<const uint32_t*>(row + 6 * (g + 8) + off_lo);
            a[m][2] = *reinterpret_cast<const uint32_t*>(row + 6 * g + off_hi);
            a[m][3] = *reinterpret_cast<const uint32_t*>(row + 6 * (g + 8) + off_hi);
        }
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
            const uint2 bf = frag_s[(s * 8 + nt) * 32 + lane];
#pragma unroll
            for (int m = 0; m < 2; ++m)
                asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
                             : "+f"(acc[m][nt][0]), "+f"(acc[m][nt][1]), "+f"(acc[m][nt][2]), "+f"(acc[m][nt][3])
                             : "r"(a[m][0]), "r"(a[m][1]), "r"(a[m][2]), "r"(a[m][3]), "r"(bf.x), "r"(bf.y));
        }
    }
    // ReLU -> bf16 -> staged tile [pixel][64] (pitch 144 B: conflict-free 4-byte writes and 16-byte reads)
#pragma unroll
    for (int m = 0; m < 2; ++m) {
        const int p0 = (warp * 2 + m) * ST_TILE + g;
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
            const __nv_bfloat162 lo = __floats2bfloat162_rn(fmaxf(acc[m][nt][0], 0.f), fmaxf(acc[m][nt][1], 0.f));
            const __nv_bfloat162 hi = __floats2bfloat162_rn(fmaxf(acc[m][nt][2], 0.f), fmaxf(acc[m][nt][3], 0.f));
            *reinterpret_cast<__nv_bfloat162*>(stage + p0 * ST_OUT_PITCH + nt * 16 + t * 4) = lo;
            *reinterpret_cast<__nv_bfloat162*>(stage + (p0 + 8) * ST_OUT_PITCH + nt * 16 + t * 4) = hi;
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < ST_TILE * ST_TILE * 8; i += 256) {
        const int px = i >> 3, ch = i & 7;
        const int oy = oy0 + px / ST_TILE, ox = ox0 + px % ST_TILE;
        if (oy < 112 && ox < 112)
            *reinterpret_cast<uint4*>(out + (((size_t)b * 112 + oy) * 112 + ox) * 64 + ch * 8) =
                *reinterpret_cast<const uint4*>(stage + px * ST_OUT_PITCH + ch * 16);
    }
}
void stem_mma(const void* pre, int B, const void* bfrag, const float* bias, void* out, cudaStream_t s) {
    ensure_dynamic_smem((const void*)(stem_mma_kernel), ST_SMEM);
    stem_mma_kernel<<<dim3(112 / ST_TILE, 112 / ST_TILE, B), 256, ST_SMEM, s>>>((const bf16*)pre, (const uint2*)bfrag, bias, (bf16*)out);
    SYNT_LAUNCH_CHECK();
}

// ---- fused classifier front-end (bf16 mode): preprocess -> conv 7x7/s2 (+BN, ReLU) -> maxpool 3x3/s2, one kernel.
// One CTA = 8x8 pooled pixels of one image = 17x17 conv pixels (one row/column of overlap with the neighbours) = a 40x40x3
// patch of the 224x224 preprocessed image, which is SAMPLED ON THE FLY from the 128x128 source (preprocess_pixel) into
// shared memory; the conv tile (ReLU'd, bf16) stays in shared memory and is pooled there: neither the 224x224 image nor the
// 112x112x64 conv output ever exists in HBM.  Out-of-plane conv pixels are written as 0, which equals the -inf padding of
// the max-pool because every window holds at least one in-plane value and all values are >= 0 after the ReLU.
// 320 threads = 10 warps x 2 m16 tiles = 320 row slots for the 289 conv pixels (row slot -> pixel is a free mapping because
// the A operand is fetched with per-thread 32-bit loads, see stem_mma_kernel).
constexpr int SF_POOL = 8, SF_CONV = 2 * SF_POOL + 1, SF_NPIX = SF_CONV * SF_CONV;          // 8, 17, 289
constexpr int SF_ROWS = 2 * SF_CONV + 6, SF_PITCH = 126;                                     // 40 patch rows (39 + the k-padding row)
constexpr int SF_PATCH_BYTES = SF_ROWS * SF_PITCH * 2;                                       // 10080
constexpr int SF_SMEM = SF_PATCH_BYTES + SF_NPIX * ST_OUT_PITCH;                             // + 41616

__global__ void __launch_bounds__(320, 2) stem_fused_kernel(const float* __restrict__ x /* [B,3,128,128] */, const uint2* __restrict__ bfrag,
                                                         const float* __restrict__ bias, bf16* __restrict__ out /* [B,56,56,64] */) {
    extern __shared__ __align__(16) uint8_t sf_smem[];
    unsigned short* patch = reinterpret_cast<unsigned short*>(sf_smem);
    uint8_t* tile = sf_smem + SF_PATCH_BYTES;
    const int b = blockIdx.z, py0 = blockIdx.y * SF_POOL, px0 = blockIdx.x * SF_POOL;
    const int cy0 = 2 * py0 - 1, cx0 = 2 * px0 - 1;                       // conv-plane origin of the tile
    {
        const float* img = x + (size_t)b * 3 * 128 * 128;
        const float sc = 128.0f / 224.0f;
        const int iy0 = 2 * cy0 - 3, ix0 = 2 * cx0 - 3;                  // preprocessed-image origin of the patch
        constexpr int PW = SF_PITCH / 3;                                 // 42 pixel slots per patch row
        for (int i = threadIdx.x; i < SF_ROWS * PW; i += 320) {
            const int r = i / PW, c = i - r * PW;
            const int iy = iy0 + r, ix = ix0 + c;
            float v3[3] = {0.f, 0.f, 0.f};
            if (iy >= 0 && iy < 224 && ix >= 0 && ix < 224) preprocess_pixel(img, 128, 128, sc, sc, iy, ix, v3);
            unsigned short* dst = patch + r * SF_PITCH + c * 3;
#pragma unroll
            for (int ch = 0; ch < 3; ++ch) dst[ch] = __bfloat16_as_ushort(__float2bfloat16_rn(v3[ch]));
        }
    }
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    // row slots of this thread: pixels p = (warp*2 + m)*16 + g (+8); slots >= 289 compute pixel 0 and are discarded
    int pbase[2][2];
#pragma unroll
    for (int m = 0; m < 2; ++m)
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            int p = (warp * 2 + m) * 16 + g + 8 * h;
            if (p >= SF_NPIX) p = 0;
            pbase[m][h] = (2 * (p / SF_CONV)) * SF_PITCH + 6 * (p % SF_CONV);
        }
    float acc[2][8][4];
#pragma unroll
    for (int m = 0; m < 2; ++m)
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
            const float b0 = __ldg(bias + nt * 8 + t * 2), b1 = __ldg(bias + nt * 8 + t * 2 + 1);
            acc[m][nt][0] = b0; acc[m][nt][1] = b1; acc[m][nt][2] = b0; acc[m][nt][3] = b1;
        }
#pragma unroll
    for (int s = 0; s < ST_STEPS; ++s) {
        const int k_lo = 16 * s, k_hi = 16 * s + 8;
        const int off_lo = (k_lo / 24) * SF_PITCH + (k_lo % 24) + 2 * t;
        const int off_hi = (k_hi / 24) * SF_PITCH + (k_hi % 24) + 2 * t;
        uint32_t a[2][4];
#pragma unroll
        for (int m = 0; m < 2; ++m) {
            a[m][0] = *reinterpret_cast<const uint32_t*>(patch + pbase[m][0] + off_lo);
            a[m][1] = *reinterpret_cast<const uint32_t*>(patch + pbase[m][1] + off_lo);
            a[m][2] = *reinterpret_cast<const uint32_t*>(patch + pbase[m][0] + off_hi);
            a[m][3] = *reinterpret_cast<const uint32_t*>(patch + pbase[m][1] + off_hi);
        }
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
            const uint2 bf = __ldg(bfrag + (s * 8 + nt) * 32 + lane);    // 22.5 KB table, L1-resident
#pragma unroll
            for (int m = 0; m < 2; ++m)
                asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
                             : "+f"(acc[m][nt][0]), "+f"(acc[m][nt][1]), "+f"(acc[m][nt][2]), "+f"(acc[m][nt][3])
                             : "r"(a[m][0]), "r"(a[m][1]), "r"(a[m][2]), "r"(a[m][3]), "r"(bf.x), "r"(bf.y));
        }
    }
    // ReLU -> bf16 -> conv tile in shared memory; pixels outside the 112x112 conv plane become 0
#pragma unroll
    for (int m = 0; m < 2; ++m)
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int p = (warp * 2 + m) * 16 + g + 8 * h;
            if (p >= SF_NPIX) continue;
            const int cy = cy0 + p / SF_CONV, cx = cx0 + p % SF_CONV;
            const bool in = cy >= 0 && cy < 112 && cx >= 0 && cx < 112;
#pragma unroll
            for (int nt = 0; nt < 8; ++nt) {
                const float v0 = in ? fmaxf(acc[m][nt][2 * h], 0.f) : 0.f, v1 = in ? fmaxf(acc[m][nt][2 * h + 1], 0.f) : 0.f;
                *reinterpret_cast<__nv_bfloat162*>(tile + p * ST_OUT_PITCH + nt * 16 + t * 4) = __floats2bfloat162_rn(v0, v1);
            }
        }
    __syncthreads();
    // maxpool 3x3 / stride 2 / pad 1 over the tile: pooled (py, px) <- conv tile rows 2py..2py+2, cols 2px..2px+2 (tile coords)
    for (int i = threadIdx.x; i < SF_POOL * SF_POOL * 8; i += 320) {
        const int pp = i >> 3, ch = i & 7, py = pp / SF_POOL, px = pp % SF_POOL;
        if (py0 + py >= 56 || px0 + px >= 56) continue;
        __nv_bfloat162 mx[4];
#pragma unroll
        for (int dy = 0; dy < 3; ++dy)
#pragma unroll
            for (int dx = 0; dx < 3; ++dx) {
                const uint4 v = *reinterpret_cast<const uint4*>(tile + ((2 * py + dy) * SF_CONV + 2 * px + dx) * ST_OUT_PITCH + ch * 16);
                const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&v);
#pragma unroll
                for (int j = 0; j < 4; ++j) mx[j] = (dy == 0 && dx == 0) ? h2[j] : __hmax2(mx[j], h2[j]);
            }
        *reinterpret_cast<uint4*>(out + (((size_t)b * 56 + py0 + py) * 56 + px0 + px) * 64 + ch * 8) = *reinterpret_cast<const uint4*>(mx);
    }
}
void stem_fused(const float* x_nchw, int B, const void* bfrag, const float* bias, void* out, cudaStream_t s) {
    ensure_dynamic_smem((const void*)(stem_fused_kernel), SF_SMEM);
    stem_fused_kernel<<<dim3(56 / SF_POOL, 56 / SF_POOL, B), 320, SF_SMEM, s>>>(x_nchw, (const uint2*)bfrag, bias, (bf16*)out);
    SYNT_LAUNCH_CHECK();
}

template <typename T>
__global__ void maxpool_kernel(const T* __restrict__ in, int H, int W, int C, int Ho, int Wo, long long nvec_total,
                               T* __restrict__ out) {
    const int nvec = C >> 3;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < nvec_total; i += (long long)gridDim.x * blockDim.x) {
        const int v = (int)(i % nvec);
        long long p = i / nvec;
        const int ox = (int)(p % Wo); p /= Wo;
        const int oy = (int)(p % Ho);
        const long long b = p / Ho;
        float m[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) m[j] = -INFINITY;
        for (int dy = 0; dy < 3; ++dy) {
            const int iy = oy * 2 - 1 + dy;
            if (iy < 0 || iy >= H) continue;
            for (int dx = 0; dx < 3; ++dx) {
                const int ix = ox * 2 - 1 + dx;
                if (ix < 0 || ix >= W) continue;
                float x[8];
                load8<T>(in + ((b * H + iy) * W + ix) * C + v * 8, x);
#pragma unroll
                for (int j = 0; j < 8; ++j) m[j] = fmaxf(m[j], x[j]);
            }
        }
        store8<T>(out + ((b * Ho + oy) * Wo + ox) * C + v * 8, m);
    }
}
void maxpool3x3s2(const void* in, int dt, int B, int H, int W, int C, void* out, cudaStream_t s) {
    const int Ho = (H + 2 - 3) / 2 + 1, Wo = (W + 2 - 3) / 2 + 1;
    const long long nv = (long long)B * Ho * Wo * (C / 8);
    const int blocks = (int)((nv + 255) / 256 < 148 * 16 ? (nv + 255) / 256 : 148 * 16);
    if (dt == DT_F32) maxpool_kernel<float><<<blocks, 256, 0, s>>>((const float*)in, H, W, C, Ho, Wo, nv, (float*)out);
    else              maxpool_kernel<bf16><<<blocks, 256, 0, s>>>((const bf16*)in, H, W, C, Ho, Wo, nv, (bf16*)out);
    SYNT_LAUNCH_CHECK();
}

template <typename T>
__global__ void __launch_bounds__(512) avgpool_fc_kernel(const T* __restrict__ in, int HW, int C, const float* __restrict__ w,
                                                         const float* __restrict__ bias, int nout, float* __restrict__ logits) {
    __shared__ float pooled[512];
    const int b = blockIdx.x, c = threadIdx.x;
    if (c < C) {
        float s = 0.f;
        for (int p = 0; p < HW; ++p) s += to_f<T>(in[((size_t)b * HW + p) * C + c]);
        pooled[c] = s / (float)HW;
    }
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int o = warp; o < nout; o += blockDim.x >> 5) {
        float s = 0.f;
        for (int i = lane; i < C; i += 32) s = fmaf(pooled[i], w[(size_t)o * C + i], s);
        s = warp_sum(s);
        if (lane == 0) logits[(size_t)b * nout + o] = s + bias[o];
    }
}
void avgpool_fc(const void* in, int dt, int B, int HW, int C, const float* w, const float* b, int nout, float* logits,
                cudaStream_t s) {
    SYNT_CHECK(C <= 512, "avgpool_fc: C <= 512");
    if (dt == DT_F32) avgpool_fc_kernel<float><<<B, 512, 0, s>>>((const float*)in, HW, C, w, b, nout, logits);
    else              avgpool_fc_kernel<bf16><<<B, 512, 0, s>>>((const bf16*)in, HW, C, w, b, nout, logits);
    SYNT_LAUNCH_CHECK();
}

// one block per (b, c) plane
__global__ void __launch_bounds__(256) intervene_kernel(const float* __restrict__ x, const float* __restrict__ mask,
                                                        const float* __restrict__ aux, int type, float noise_std, int blur_k,
                                                        int C, int H, int W, float* __restrict__ out,
                                                        float* __restrict__ interv_out) {
    __shared__ float red[256];
    __shared__ float mean_s;
    const int plane = blockIdx.x, b = plane / C, HW = H * W;
    const float* xp = x + (size_t)plane * HW;
    const float* mp = mask + (size_t)b * HW;
    float* op = out + (size_t)plane * HW;
    if (type == 1) {
        float s = 0.f;
        for (int i = threadIdx.x; i < HW; i += 256) s += xp[i];
        red[threadIdx.x] = s;
        __syncthreads();
        for (int o = 128; o > 0; o >>= 1) { if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o]; __syncthreads(); }
        if (threadIdx.x == 0) mean_s = red[0] / (float)HW;
        __syncthreads();
    }
    for (int i = threadIdx.x; i < HW; i += 256) {
        float iv;
        if (type == 0) iv = 0.f;
        else if (type == 1) iv = mean_s;
        else if (type == 2) {
            const int y = i / W, xx = i % W, r = blur_k / 2;
            float s = 0.f;
            for (int dy = -r; dy <= r; ++dy)
                for (int dx = -r; dx <= r; ++dx) {
                    const int yy = y + dy, xc = xx + dx;
                    if (yy >= 0 && yy < H && xc >= 0 && xc < W) s += xp[yy * W + xc];
                }
            iv = s / (float)(blur_k * blur_k);             // avg_pool2d count_include_pad=True
        } else if (type == 3) iv = aux[(size_t)plane * HW + i] * noise_std;
        else iv = aux[(size_t)plane * HW + i];
        const float m = mp[i];
        const float v = xp[i] * (1.f - m) + iv * m;
        op[i] = fminf(fmaxf(v, -1.f), 1.f);
        if (interv_out) interv_out[(size_t)plane * HW + i] = iv;
    }
}
void intervene_blend(const float* x, const float* mask, const float* aux, int type, float noise_std, int blur_k, int B, int C,
                     int H, int W, float* out, float* interv_out, cudaStream_t s) {
    SYNT_CHECK(type >= 0 && type <= 4, "intervene: unknown type");
    SYNT_CHECK(type < 3 || aux != nullptr, "intervene: aux tensor required");
    SYNT_CHECK(blur_k >= 1 && (blur_k & 1) && blur_k <= 63, "intervene: blur kernel must be odd, 1..63");
    intervene_kernel<<<B * C, 256, 0, s>>>(x, mask, aux, type, noise_std, blur_k, C, H, W, out, interv_out);
    SYNT_LAUNCH_CHECK();
}

__global__ void patch_mask_kernel(const float* __restrict__ x, const unsigned char* __restrict__ pm, int C, int H, int W,
                                  int patch, long long n, float* __restrict__ out) {
    const int pw = W / patch, ph = H / patch;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const int xx = (int)(i % W), y = (int)((i / W) % H);
        const int c = (int)((i / ((long long)W * H)) % C);
        const long long k = i / ((long long)W * H * C);
        const bool keep = pm[(k * ph + y / patch) * pw + xx / patch] != 0;
        out[i] = keep ? x[((size_t)c * H + y) * W + xx] : 0.f;
    }
}
void patch_mask_apply(const float* x, const unsigned char* pm, int n_masks, int C, int H, int W, int patch, float* out,
                      cudaStream_t s) {
    const long long n = (long long)n_masks * C * H * W;
    const int blocks = (int)((n + 255) / 256 < 148 * 16 ? (n + 255) / 256 : 148 * 16);
    patch_mask_kernel<<<blocks, 256, 0, s>>>(x, pm, C, H, W, patch, n, out);
    SYNT_LAUNCH_CHECK();
}

// =============================================================== plan ================
struct RParam { std::string name; long long numel, offset; };
struct RManifest {
    std::vector<RParam> params; long long total = 0;
    void add(const std::string& n, long long numel) { params.push_back({n, numel, total}); total += numel; }
    void bn(const std::string& p, int c) { add(p + ".weight", c); add(p + ".bias", c); add(p + ".running_mean", c); add(p + ".running_var", c); }
    long long find(const std::string& n) const {
        for (auto& s : params) if (s.name == n) return s.offset;
        throw Error(-1, "resnet manifest: no parameter named " + n);
    }
};
static const int kStageCh[4] = {64, 128, 256, 512};
static RManifest build_rmanifest(int num_classes) {
    RManifest m;
    m.add("conv1.weight", 64 * 3 * 49); m.bn("bn1", 64);
    int cin = 64;
    for (int l = 0; l < 4; ++l) {
        const int c = kStageCh[l];
        for (int j = 0; j < 2; ++j) {
            const std::string p = "layer" + std::to_string(l + 1) + "." + std::to_string(j);
            const int ci = j == 0 ? cin : c;
            m.add(p + ".conv1.weight", (long long)c * ci * 9); m.bn(p + ".bn1", c);
            m.add(p + ".conv2.weight", (long long)c * c * 9); m.bn(p + ".bn2", c);
            if (j == 0 && l > 0) { m.add(p + ".downsample.0.weight", (long long)c * ci); m.bn(p + ".downsample.1", c); }
        }
        cin = c;
    }
    m.add("fc.weight", (long long)num_classes * 512); m.add("fc.bias", num_classes);
    return m;
}
static const RManifest& rmanifest(int nc = 7) {
    static RManifest m7 = build_rmanifest(7);
    static RManifest m8 = build_rmanifest(8);
    if (nc == 8) return m8;
    SYNT_CHECK(nc == 7, "resnet18 head must have 7 (xai_integration.py:79) or 8 (XAI.py:490) outputs");
    return m7;
}

struct RDev { void* p = nullptr; ~RDev() { if (p) cudaFree(p); } };
using RPtr = std::shared_ptr<RDev>;
static RPtr r_upload(const void* h, size_t bytes) {
    auto b = std::make_shared<RDev>();
    SYNT_CUDA(cudaMalloc(&b->p, bytes));
    SYNT_CUDA(cudaMemcpy(b->p, h, bytes, cudaMemcpyHostToDevice));
    return b;
}
static uint16_t r_f2bf(float f) { uint32_t u; memcpy(&u, &f, 4); return (uint16_t)((u + 0x7FFFu + ((u >> 16) & 1u)) >> 16); }

struct RConv {                     // BN-folded conv: weight K-major [cout][k*k*cin (+ csc)], bias [cout]
    int cin = 0, cout = 0, k = 3, stride = 1, csc = 0;
    RPtr w; RPtr b; bool bf = false;
};

// fold eval-mode BN into (w, b): w' = w * g/sqrt(var+eps), b' = beta - mean*g/sqrt(var+eps)
static void fold_bn(const float* P, const RManifest& m, const std::string& bn, int cout, std::vector<float>& scale,
                    std::vector<float>& shift) {
    const float* g = P + m.find(bn + ".weight"); const float* be = P + m.find(bn + ".bias");
    const float* mu = P + m.find(bn + ".running_mean"); const float* var = P + m.find(bn + ".running_var");
    scale.resize(cout); shift.resize(cout);
    for (int i = 0; i < cout; ++i) {
        const float s = g[i] / std::sqrt(var[i] + 1e-5f);
        scale[i] = s; shift[i] = be[i] - mu[i] * s;
    }
}
static RConv make_rconv(const float* P, const RManifest& m, const std::string& conv, const std::string& bn, int cin, int cout,
                        int k, int stride, const std::string& ds_conv, const std::string& ds_bn, int ds_cin, bool bf) {
    RConv c; c.cin = cin; c.cout = cout; c.k = k; c.stride = stride; c.csc = ds_conv.empty() ? 0 : ds_cin; c.bf = bf;
    const int taps = k * k, ktot = taps * cin + c.csc;
    std::vector<float> sc, sh, sc2, sh2;
    fold_bn(P, m, bn, cout, sc, sh);
    if (c.csc) fold_bn(P, m, ds_bn, cout, sc2, sh2);
    const float* w = P + m.find(conv + ".weight");
    const float* wd = c.csc ? P + m.find(ds_conv + ".weight") : nullptr;
    std::vector<float> pk((size_t)cout * ktot), bias(cout);
    for (int n = 0; n < cout; ++n) {
        float* row = pk.data() + (size_t)n * ktot;
        for (int t = 0; t < taps; ++t)
            for (int ci = 0; ci < cin; ++ci) row[t * cin + ci] = w[((size_t)n * cin + ci) * taps + t] * sc[n];
        for (int ci = 0; ci < c.csc; ++ci) row[taps * cin + ci] = wd[(size_t)n * c.csc + ci] * sc2[n];
        bias[n] = sh[n] + (c.csc ? sh2[n] : 0.f);
    }
    if (bf) {
        std::vector<uint16_t> h(pk.size());
        for (size_t i = 0; i < pk.size(); ++i) h[i] = r_f2bf(pk[i]);
        c.w = r_upload(h.data(), h.size() * 2);
    } else {
        c.w = r_upload(pk.data(), pk.size() * 4);
    }
    c.b = r_upload(bias.data(), bias.size() * 4);
    return c;
}

}  // namespace synt

using namespace synt;

struct synt_resnet18 {
    int dt = DT_BF16, num_classes = 7; bool use_tc = true;
    bool use_v2 = true;                      // conv_tc2 where its shape rules allow (SYNT_RESNET_V2=0: conv_tc everywhere)
    bool fuse_front = true;                  // preprocess + stem + maxpool in one kernel (SYNT_RESNET_FUSE_FRONT=0: three kernels)
    RConv stem;                              // 7x7 s2 on the fp32-FMA kernel (fp32 mode)
    RConv stem_tc;                           // the same stem as a 1x1 conv over the im2col'd input (K 147 -> 192), bf16 tcgen05
    RPtr stem_frag;                          // stem weights as mma.sync B fragments (mma.sync stem kernels)
    RPtr stem_taps;                          // stem weights as 16 space-to-depth tap tiles (tcgen05 front end, default in bf16 mode)
    RConv c1[4][2], c2[4][2];
    // data-gradient convolutions of the same layers (built on first use by ensure_bwd): d2 = dgrad of conv2 (c -> c),
    // d1 = dgrad of conv1 (c -> block input channels; the stride-2 blocks carry the 1x1 downsample dgrad as shortcut segment)
    RConv d1[4][2], d2[4][2];
    RPtr zero_bias; bool bwd_ready = false;
    RPtr fc_w, fc_b;
    Pool pool;
    long long launches = 0;
    std::string tap_name; float* tap_out = nullptr; long long tap_cap = 0; int tap_C = 0, tap_H = 0, tap_W = 0; bool tap_hit = false;
};

namespace synt {

void stem_im2col(const void* pre, int B, void* out, cudaStream_t s);
void stem_mma(const void* pre, int B, const void* bfrag, const float* bias, void* out, cudaStream_t s);
void stem_fused(const float* x_nchw, int B, const void* bfrag, const float* bias, void* out, cudaStream_t s);
// tcgen05 front end (stem_tc.cu)
void stem_tc(const float* x_nchw, int B, const void* w_taps, const float* bias, void* out, cudaStream_t s);
void stem_tc_pack_weights(const uint16_t* w_k192, uint16_t* out);
int stem_tc_weight_bytes();

struct RFwd {
    synt_resnet18* r; cudaStream_t s; int B;
    cudaEvent_t* ev = nullptr;               // measurement hook (synt_resnet18_profile): begin / front end / body / head
    void mark(int i) { if (ev) cudaEventRecord(ev[i], s); }
    size_t esz() const { return dtype_size(r->dt); }
    void* make(int H, int W, int C) { return r->pool.alloc((size_t)B * H * W * C * esz()); }
    void tap(const std::string& name, void* p, int H, int W, int C) {
        if (!r->tap_out || name != r->tap_name) return;
        SYNT_CHECK((long long)B * H * W * C <= r->tap_cap, "debug tap buffer too small");
        nhwc_to_nchw_f32(p, r->dt, B, H * W, C, r->tap_out, s);
        r->tap_C = C; r->tap_H = H; r->tap_W = W; r->tap_hit = true;
    }
    void conv(const RConv& c, const void* in, int H, int W, const void* sc, int sc_stride, const void* residual, int relu,
              void* out, int Ho, int Wo, bool allow_v2 = true) {
        ConvArgs a; a.in = in; a.B = B; a.H = H; a.W = W; a.Cin = c.cin; a.KH = a.KW = c.k; a.stride = c.stride;
        a.pad = c.k / 2; a.Ho = Ho; a.Wo = Wo; a.Cout = c.cout;
        if (c.csc) { a.sc0 = sc; a.sc0_C = c.csc; a.sc_stride = sc_stride; }
        a.weight = c.w->p; a.bias = (const float*)c.b->p; a.residual = residual; a.relu = relu; a.out = out;
        // stride-1 3x3 convs on the 56x56 / 28x28 planes: the persistent halo-tile kernel (ragged tile grid); the rest
        // (stride 2, strided 1x1 shortcut segments, 14x14 and 7x7 planes): one TMA box per tap
        if (c.bf && r->use_v2 && allow_v2 && conv_tc2_supported(a)) conv_tc2(a, s);
        else if (c.bf) conv_tc(a, s);
        else conv_simt(a, r->dt, s);
        ++r->launches;
    }
    void run(const float* x_nchw, float* logits) {
        void* cur = nullptr;
        mark(0);
        if (r->use_tc && r->stem_frag && r->fuse_front && !r->tap_out) {
            // preprocess + stem + maxpool in one kernel (the debug taps "preprocess" / "relu" / "maxpool" use the unfused path)
            cur = make(56, 56, 64);
            if (r->stem_taps) stem_tc(x_nchw, B, r->stem_taps->p, (const float*)r->stem_tc.b->p, cur, s);
            else stem_fused(x_nchw, B, r->stem_frag->p, (const float*)r->stem_tc.b->p, cur, s);
            ++r->launches;
        } else {
            void* pre = make(224, 224, 3);
            classifier_preprocess(x_nchw, B, 128, 128, 224, 224, pre, 3, r->dt, s);
            tap("preprocess", pre, 224, 224, 3);
            void* c1 = make(112, 112, 64);
            if (r->use_tc && r->stem_frag) {                     // fused 7x7/s2 stem on mma.sync, no im2col tensor
                stem_mma(pre, B, r->stem_frag->p, (const float*)r->stem_tc.b->p, c1, s);
                ++r->launches;
            } else if (r->use_tc) {
                void* col = make(112, 112, 192);
                stem_im2col(pre, B, col, s);
                ++r->launches;
                conv(r->stem_tc, col, 112, 112, nullptr, 1, nullptr, 1, c1, 112, 112);
                r->pool.release(col);
            } else {
                conv(r->stem, pre, 224, 224, nullptr, 1, nullptr, 1, c1, 112, 112);
            }
            r->pool.release(pre);
            tap("relu", c1, 112, 112, 64);
            cur = make(56, 56, 64);
            maxpool3x3s2(c1, r->dt, B, 112, 112, 64, cur, s);
            r->pool.release(c1);
            tap("maxpool", cur, 56, 56, 64);
            r->launches += 2;
        }
        mark(1);
        int H = 56;
        for (int l = 0; l < 4; ++l) {
            const int c = kStageCh[l];
            for (int j = 0; j < 2; ++j) {
                const int stride = (j == 0 && l > 0) ? 2 : 1;
                const int Ho = H / stride;
                void* h1 = make(Ho, Ho, c);
                conv(r->c1[l][j], cur, H, H, nullptr, 1, nullptr, 1, h1, Ho, Ho);
                void* o = make(Ho, Ho, c);
                if (r->c2[l][j].csc) conv(r->c2[l][j], h1, Ho, Ho, cur, stride, nullptr, 1, o, Ho, Ho);
                else                 conv(r->c2[l][j], h1, Ho, Ho, nullptr, 1, cur, 1, o, Ho, Ho);
                r->pool.release(h1); r->pool.release(cur);
                cur = o; H = Ho;
                tap("layer" + std::to_string(l + 1) + "." + std::to_string(j), cur, H, H, c);
            }
        }
        mark(2);
        avgpool_fc(cur, r->dt, B, H * H, 512, (const float*)r->fc_w->p, (const float*)r->fc_b->p, r->num_classes, logits, s);
        ++r->launches;
        r->pool.release(cur);
        mark(3);
    }
};

// ---- input gradient: forward with every activation kept, then the adjoint chain (kernels: resnet_grad.cu) --------------
static RPtr r_alloc(size_t bytes) {
    auto b = std::make_shared<RDev>();
    SYNT_CUDA(cudaMalloc(&b->p, bytes));
    return b;
}
static void ensure_bwd(synt_resnet18* r, cudaStream_t s) {
    if (r->bwd_ready) return;
    const bool bf = r->use_tc;
    const size_t es = bf ? 2 : 4;
    std::vector<float> z(512, 0.f);
    r->zero_bias = r_upload(z.data(), z.size() * 4);
    int cin = 64;
    for (int l = 0; l < 4; ++l) {
        const int c = kStageCh[l];
        for (int j = 0; j < 2; ++j) {
            const bool down = j == 0 && l > 0;
            const int ci = j == 0 ? cin : c;
            const RConv& f1 = r->c1[l][j];
            const RConv& f2 = r->c2[l][j];
            RConv& d2 = r->d2[l][j];                           // gradient at conv2's output [c] -> conv2's input [c]
            d2.cin = c; d2.cout = c; d2.k = 3; d2.stride = 1; d2.csc = 0; d2.bf = bf; d2.b = r->zero_bias;
            d2.w = r_alloc((size_t)c * 9 * c * es);
            dgrad_weights(f2.w->p, 9 * c + f2.csc, 0, c, c, 9, bf, d2.w->p, 9 * c, 0, s);
            RConv& d1 = r->d1[l][j];                           // gradient at conv1's output [c] -> block input [ci]
            d1.cin = c; d1.cout = ci; d1.k = 3; d1.stride = 1; d1.csc = down ? c : 0; d1.bf = bf; d1.b = r->zero_bias;
            const int ld = 9 * c + d1.csc;
            d1.w = r_alloc((size_t)ci * ld * es);
            dgrad_weights(f1.w->p, 9 * ci, 0, ci, c, 9, bf, d1.w->p, ld, 0, s);
            if (down) dgrad_weights(f2.w->p, 9 * c + f2.csc, 9 * c, ci, c, 1, bf, d1.w->p, ld, 9 * c, s);
        }
        cin = c;
    }
    SYNT_CUDA(cudaStreamSynchronize(s));
    r->bwd_ready = true;
}

struct RGrad {
    synt_resnet18* r; cudaStream_t s; int B;
    void gtap(const std::string& name, const void* p, int dt, int H, int W, int C) {
        if (!r->tap_out || name != r->tap_name) return;
        SYNT_CHECK((long long)B * H * W * C <= r->tap_cap, "debug tap buffer too small");
        nhwc_to_nchw_f32(p, dt, B, H * W, C, r->tap_out, s);
        r->tap_C = C; r->tap_H = H; r->tap_W = W; r->tap_hit = true;
    }
    // score[b] = log(softmax(logits)[target] + 1e-8) (nullable), dx[b] = d score[b] / d x[b]  (fp32 NCHW like x)
    void run(const float* x_nchw, int target, float* score, float* dx) {
        ensure_bwd(r, s);
        RFwd f{r, s, B};
        const int dt = r->dt;
        Pool& pool = r->pool;
        // ---------------- forward, activations kept ----------------
        void* pre = f.make(224, 224, 3);
        classifier_preprocess(x_nchw, B, 128, 128, 224, 224, pre, 3, dt, s);
        void* a1 = f.make(112, 112, 64);                            // stem output (post-ReLU)
        if (r->use_tc && r->stem_frag) {
            stem_mma(pre, B, r->stem_frag->p, (const float*)r->stem_tc.b->p, a1, s);
        } else if (r->use_tc) {
            void* col = f.make(112, 112, 192);
            stem_im2col(pre, B, col, s);
            f.conv(r->stem_tc, col, 112, 112, nullptr, 1, nullptr, 1, a1, 112, 112);
            pool.release(col);
        } else {
            f.conv(r->stem, pre, 224, 224, nullptr, 1, nullptr, 1, a1, 112, 112);
        }
        pool.release(pre);
        void* x0 = f.make(56, 56, 64);
        unsigned char* pool_idx = (unsigned char*)pool.alloc((size_t)B * 56 * 56 * 64);
        maxpool3x3s2_idx(a1, dt, B, 112, 112, 64, x0, pool_idx, s);
        void* h1[4][2]; void* o[4][2];
        const void* cur = x0;
        int H = 56;
        for (int l = 0; l < 4; ++l) {
            const int c = kStageCh[l];
            for (int j = 0; j < 2; ++j) {
                const int stride = (j == 0 && l > 0) ? 2 : 1;
                const int Ho = H / stride;
                h1[l][j] = f.make(Ho, Ho, c);
                f.conv(r->c1[l][j], cur, H, H, nullptr, 1, nullptr, 1, h1[l][j], Ho, Ho);
                o[l][j] = f.make(Ho, Ho, c);
                if (r->c2[l][j].csc) f.conv(r->c2[l][j], h1[l][j], Ho, Ho, cur, stride, nullptr, 1, o[l][j], Ho, Ho);
                else                 f.conv(r->c2[l][j], h1[l][j], Ho, Ho, nullptr, 1, cur, 1, o[l][j], Ho, Ho);
                cur = o[l][j]; H = Ho;
            }
        }
        const int nc = r->num_classes;
        float* logits = (float*)pool.alloc((size_t)B * nc * 4);
        float* dlog = (float*)pool.alloc((size_t)B * nc * 4);
        avgpool_fc(cur, dt, B, H * H, 512, (const float*)r->fc_w->p, (const float*)r->fc_b->p, nc, logits, s);
        score_grad(logits, B, nc, target, score, dlog, s);
        // ---------------- backward ----------------
        // g is always the gradient at a block's PRE-ReLU output (the ReLU mask of the tensor it arrives at is applied)
        void* g = f.make(7, 7, 512);
        avgpool_fc_bwd(dlog, (const float*)r->fc_w->p, o[3][1], dt, B, 49, 512, nc, g, s);
        pool.release(logits); pool.release(dlog);
        r->launches += 5;
        for (int l = 3; l >= 0; --l) {
            const int c = kStageCh[l];
            const int Ho = 56 >> l;
            for (int j = 1; j >= 0; --j) {
                const bool down = j == 0 && l > 0;
                const int ci = down ? kStageCh[l - 1] : c;
                const int Hin = down ? 2 * Ho : Ho;
                gtap("grad:layer" + std::to_string(l + 1) + "." + std::to_string(j), g, dt, Ho, Ho, c);
                void* t = f.make(Ho, Ho, c);
                f.conv(r->d2[l][j], g, Ho, Ho, nullptr, 1, nullptr, 0, t, Ho, Ho);
                void* dxb = f.make(Hin, Hin, ci);
                if (!down) {
                    relu_mask(t, h1[l][j], dt, (long long)B * Ho * Ho * c, t, s);
                    f.conv(r->d1[l][j], t, Ho, Ho, nullptr, 1, /*identity shortcut*/ g, 0, dxb, Ho, Ho);
                } else {
                    // stride-2 block: conv1 (3x3/s2) and the 1x1/s2 downsample both read the block input; their data
                    // gradients are stride-1 convolutions over the zero-inserted gradient planes, fused as main + shortcut
                    void* tz = f.make(Hin, Hin, c);
                    void* gz = f.make(Hin, Hin, c);
                    zero_insert2x(t, h1[l][j], dt, B, Ho, Ho, c, tz, s);
                    zero_insert2x(g, nullptr, dt, B, Ho, Ho, c, gz, s);
                    f.conv(r->d1[l][j], tz, Hin, Hin, gz, 1, nullptr, 0, dxb, Hin, Hin, /*allow_v2=*/false);
                    pool.release(tz); pool.release(gz);
                    ++r->launches;
                }
                ++r->launches;
                pool.release(t); pool.release(g);
                pool.release(h1[l][j]); pool.release(o[l][j]);
                // the block input is the previous block's ReLU output (or, for layer1.0, the max-pool output: no mask)
                const void* xin = (l == 0 && j == 0) ? nullptr : (j == 1 ? o[l][0] : o[l - 1][1]);
                if (xin) { relu_mask(dxb, xin, dt, (long long)B * Hin * Hin * ci, dxb, s); ++r->launches; }
                g = dxb;
            }
        }
        gtap("grad:maxpool", g, dt, 56, 56, 64);
        void* da1 = f.make(112, 112, 64);
        maxpool3x3s2_bwd(g, pool_idx, a1, dt, B, 112, 112, 64, da1, s);
        gtap("grad:relu", da1, dt, 112, 112, 64);
        pool.release(g); pool.release(a1); pool.release(x0); pool.release(pool_idx);
        float* dpre = (float*)pool.alloc((size_t)B * 224 * 224 * 3 * 4);
        stem_dgrad(da1, dt, B, (const float*)r->stem.w->p, dpre, s);
        gtap("grad:preprocess", dpre, DT_F32, 224, 224, 3);
        pool.release(da1);
        classifier_preprocess_bwd(dpre, x_nchw, B, 128, 128, 224, 224, dx, s);
        pool.release(dpre);
        r->launches += 3;
    }
};

}  // namespace synt

#define SYNT_TRY try {
#define SYNT_CATCH                                                                    \
    } catch (const synt::Error& e) { synt::g_last_error = e.what(); return e.code;    \
    } catch (const std::exception& e) { synt::g_last_error = e.what(); return -1; }   \
    return 0;

extern "C" {

int synt_resnet18_num_params(void) { return (int)rmanifest(7).params.size(); }
int synt_resnet18_param_info(int num_classes, int i, char* name, int cap, long long* numel, long long* offset) {
    SYNT_TRY
    const RManifest& m = rmanifest(num_classes);
    SYNT_CHECK(i >= 0 && i < (int)m.params.size(), "param index out of range");
    if (name && cap > 0) { strncpy(name, m.params[i].name.c_str(), cap - 1); name[cap - 1] = 0; }
    if (numel) *numel = m.params[i].numel;
    if (offset) *offset = m.params[i].offset;
    SYNT_CATCH
}

int synt_resnet18_create(const float* P, long long n_params, int num_classes, int dtype, synt_resnet18_t** out) {
    SYNT_TRY
    SYNT_CHECK(P && out, "null argument");
    const RManifest& m = rmanifest(num_classes);
    SYNT_CHECK(n_params == m.total, "resnet18 parameter blob has the wrong length");
    SYNT_CHECK(dtype == DT_F32 || dtype == DT_BF16, "bad dtype");
    int ndev = 0;
    SYNT_CUDA(cudaGetDeviceCount(&ndev));
    SYNT_CHECK(ndev > 0, "no CUDA device: this library has no CPU fallback");
    std::unique_ptr<synt_resnet18> r(new synt_resnet18());
    r->dt = dtype; r->num_classes = num_classes;
    const char* force = getenv("SYNT_FORCE_SIMT");
    r->use_tc = dtype == DT_BF16 && !(force && force[0] == '1');
    const bool bf = r->use_tc;
    { const char* v2 = getenv("SYNT_RESNET_V2"); r->use_v2 = !(v2 && v2[0] == '0'); }
    { const char* ff = getenv("SYNT_RESNET_FUSE_FRONT"); r->fuse_front = !(ff && ff[0] == '0'); }
    r->stem = make_rconv(P, m, "conv1", "bn1", 3, 64, 7, 2, "", "", 0, false);
    if (bf) {                                               // [64][3][7][7] -> K-major [64][ky*24 + kx*3 + c], zero-padded to 192, bf16
        std::vector<float> sc, sh;
        fold_bn(P, m, "bn1", 64, sc, sh);
        const float* w = P + m.find("conv1.weight");        // [64][3][7][7]
        std::vector<uint16_t> pk((size_t)64 * 192, 0);
        for (int n = 0; n < 64; ++n)
            for (int ky = 0; ky < 7; ++ky)
                for (int kx = 0; kx < 7; ++kx)
                    for (int c = 0; c < 3; ++c)
                        pk[(size_t)n * 192 + ky * 24 + kx * 3 + c] = r_f2bf(w[((size_t)n * 3 + c) * 49 + ky * 7 + kx] * sc[n]);
        r->stem_tc.cin = 192; r->stem_tc.cout = 64; r->stem_tc.k = 1; r->stem_tc.stride = 1; r->stem_tc.bf = true;
        r->stem_tc.w = r_upload(pk.data(), pk.size() * 2);
        r->stem_tc.b = r_upload(sh.data(), sh.size() * 4);
        const char* sm = getenv("SYNT_STEM_IM2COL");
        if (!(sm && sm[0] == '1')) {                         // per-lane mma.sync B fragments [11 k-steps][8 n-tiles][32 lanes]
            std::vector<uint32_t> fr((size_t)11 * 8 * 32 * 2);
            auto wv = [&](int k, int n) -> uint32_t { return pk[(size_t)n * 192 + k]; };
            for (int st = 0; st < 11; ++st)
                for (int nt = 0; nt < 8; ++nt)
                    for (int lane = 0; lane < 32; ++lane) {
                        const int g = lane >> 2, t = lane & 3, k0 = st * 16 + t * 2, n = nt * 8 + g;
                        uint32_t* o = fr.data() + ((size_t)(st * 8 + nt) * 32 + lane) * 2;
                        o[0] = wv(k0, n) | (wv(k0 + 1, n) << 16);
                        o[1] = wv(k0 + 8, n) | (wv(k0 + 9, n) << 16);
                    }
            r->stem_frag = r_upload(fr.data(), fr.size() * 4);
            const char* st = getenv("SYNT_STEM_TC");         // =0: the mma.sync fused front end instead of the tcgen05 one
            if (!(st && st[0] == '0')) {
                std::vector<uint16_t> taps((size_t)stem_tc_weight_bytes() / 2);
                stem_tc_pack_weights(pk.data(), taps.data());
                r->stem_taps = r_upload(taps.data(), taps.size() * 2);
            }
        }
    }
    int cin = 64;
    for (int l = 0; l < 4; ++l) {
        const int c = kStageCh[l];
        for (int j = 0; j < 2; ++j) {
            const std::string p = "layer" + std::to_string(l + 1) + "." + std::to_string(j);
            const int ci = j == 0 ? cin : c, stride = (j == 0 && l > 0) ? 2 : 1;
            r->c1[l][j] = make_rconv(P, m, p + ".conv1", p + ".bn1", ci, c, 3, stride, "", "", 0, bf);
            if (j == 0 && l > 0)
                r->c2[l][j] = make_rconv(P, m, p + ".conv2", p + ".bn2", c, c, 3, 1, p + ".downsample.0", p + ".downsample.1", ci, bf);
            else
                r->c2[l][j] = make_rconv(P, m, p + ".conv2", p + ".bn2", c, c, 3, 1, "", "", 0, bf);
        }
        cin = c;
    }
    r->fc_w = r_upload(P + m.find("fc.weight"), (size_t)num_classes * 512 * 4);
    r->fc_b = r_upload(P + m.find("fc.bias"), (size_t)num_classes * 4);
    *out = r.release();
    SYNT_CATCH
}
int synt_resnet18_destroy(synt_resnet18_t* h) { delete h; return 0; }

int synt_resnet18_logits(synt_resnet18_t* h, const float* x, int B, float* logits, void* stream) {
    SYNT_TRY
    SYNT_CHECK(h && x && logits && B > 0, "bad argument");
    const size_t img = (size_t)3 * 128 * 128;
    // micro-batch: bounds the workspace.  Fused front-end: ~1.3 MB per image (largest tensors: 56x56x64 bf16) -> 512 images
    // per pass keep the 14x14 / 7x7 layers' grids full; unfused paths materialise the 224x224 image / the stem output
    const bool fused = h->use_tc && h->stem_frag && h->fuse_front && !h->tap_out;
    static const int mb_env = [] { const char* e = getenv("SYNT_RESNET_MB"); return e ? atoi(e) : 0; }();
    const int mb = mb_env > 0 ? mb_env : (fused ? 512 : 128);
    for (int b0 = 0; b0 < B; b0 += mb) {
        RFwd f{h, (cudaStream_t)stream, B - b0 < mb ? B - b0 : mb};
        f.run(x + b0 * img, logits + (size_t)b0 * h->num_classes);
    }
    SYNT_CATCH
}
// measurement hook (bench.py roofline_hbm): one warm forward of B <= 512 images with CUDA events between the front end
// (preprocess + 7x7 stem + max-pool), the 16 body convolutions and the pool + FC head; ms_out[3]
int synt_resnet18_profile(synt_resnet18_t* h, const float* x, int B, float* logits, double* ms_out, void* stream) {
    SYNT_TRY
    SYNT_CHECK(h && x && logits && ms_out && B > 0 && B <= 512, "bad argument (B <= 512)");
    cudaStream_t s = (cudaStream_t)stream;
    { RFwd warm{h, s, B}; warm.run(x, logits); }
    cudaEvent_t ev[4];
    for (auto& e : ev) SYNT_CUDA(cudaEventCreate(&e));
    RFwd f{h, s, B};
    f.ev = ev;
    f.run(x, logits);
    SYNT_CUDA(cudaStreamSynchronize(s));
    for (int i = 0; i < 3; ++i) { float ms = 0.f; SYNT_CUDA(cudaEventElapsedTime(&ms, ev[i], ev[i + 1])); ms_out[i] = ms; }
    for (auto& e : ev) cudaEventDestroy(e);
    SYNT_CATCH
}
int synt_resnet18_logits_host(synt_resnet18_t* h, const float* x_host, int B, float* logits_host) {
    SYNT_TRY
    SYNT_CHECK(h && x_host && logits_host && B > 0, "bad argument");
    const size_t n = (size_t)B * 3 * 128 * 128;
    float* x = (float*)h->pool.alloc(n * 4);
    float* lg = (float*)h->pool.alloc((size_t)B * h->num_classes * 4);
    SYNT_CUDA(cudaMemcpyAsync(x, x_host, n * 4, cudaMemcpyHostToDevice, 0));
    int rc = synt_resnet18_logits(h, x, B, lg, nullptr);
    if (rc == 0) {
        SYNT_CUDA(cudaMemcpyAsync(logits_host, lg, (size_t)B * h->num_classes * 4, cudaMemcpyDeviceToHost, 0));
        SYNT_CUDA(cudaStreamSynchronize(0));
    }
    h->pool.release(x); h->pool.release(lg);
    if (rc != 0) return rc;
    SYNT_CATCH
}
int synt_resnet18_debug(synt_resnet18_t* h, const float* x, int B, const char* tap, float* out, long long cap, int* C,
                        int* H, int* W, void* stream) {
    SYNT_TRY
    SYNT_CHECK(h && x && tap && out && B > 0 && B <= 64, "bad argument");
    h->tap_name = tap; h->tap_out = out; h->tap_cap = cap; h->tap_hit = false;
    float* lg = (float*)h->pool.alloc((size_t)B * h->num_classes * 4);
    RFwd f{h, (cudaStream_t)stream, B};
    f.run(x, lg);
    h->pool.release(lg);
    h->tap_out = nullptr;
    SYNT_CHECK(h->tap_hit, std::string("unknown debug tap: ") + tap);
    if (C) *C = h->tap_C;
    if (H) *H = h->tap_H;
    if (W) *W = h->tap_W;
    SYNT_CATCH
}
long long synt_resnet18_launch_count(synt_resnet18_t* h) { return h ? h->launches : 0; }

int synt_resnet18_score_grad(synt_resnet18_t* h, const float* x, int B, int target_class, float* score, float* grad,
                             void* stream) {
    SYNT_TRY
    SYNT_CHECK(h && x && grad && B > 0, "bad argument");
    SYNT_CHECK(target_class >= 0 && target_class < h->num_classes, "target class out of range");
    const size_t img = (size_t)3 * 128 * 128;
    // every activation of the micro-batch stays resident until its adjoint has run: ~5.3 MB (bf16) per image
    const int mb = h->dt == DT_BF16 ? 64 : 32;
    for (int b0 = 0; b0 < B; b0 += mb) {
        RGrad g{h, (cudaStream_t)stream, B - b0 < mb ? B - b0 : mb};
        g.run(x + b0 * img, target_class, score ? score + b0 : nullptr, grad + b0 * img);
    }
    SYNT_CATCH
}
int synt_resnet18_grad_debug(synt_resnet18_t* h, const float* x, int B, int target_class, const char* tap, float* out,
                             long long cap, int* C, int* H, int* W, void* stream) {
    SYNT_TRY
    SYNT_CHECK(h && x && tap && out && B > 0 && B <= 32, "bad argument");
    SYNT_CHECK(target_class >= 0 && target_class < h->num_classes, "target class out of range");
    h->tap_name = tap; h->tap_out = out; h->tap_cap = cap; h->tap_hit = false;
    float* dx = (float*)h->pool.alloc((size_t)B * 3 * 128 * 128 * 4);
    RGrad g{h, (cudaStream_t)stream, B};
    try { g.run(x, target_class, nullptr, dx); } catch (...) { h->tap_out = nullptr; h->pool.release(dx); throw; }
    h->pool.release(dx);
    h->tap_out = nullptr;
    SYNT_CHECK(h->tap_hit, std::string("unknown gradient tap: ") + tap);
    if (C) *C = h->tap_C;
    if (H) *H = h->tap_H;
    if (W) *W = h->tap_W;
    SYNT_CATCH
}
int synt_ig_interpolate(const float* x, const float* baseline, int n_steps, long long per_image, float* out, void* stream) {
    SYNT_TRY
    SYNT_CHECK(x && baseline && out && n_steps > 0 && per_image > 0, "bad argument");
    ig_interpolate(x, baseline, n_steps, per_image, out, (cudaStream_t)stream);
    SYNT_CATCH
}
int synt_ig_reduce(const float* grads, const float* x, const float* baseline, int n_steps, long long per_image, float* out,
                   void* stream) {
    SYNT_TRY
    SYNT_CHECK(grads && x && baseline && out && n_steps > 0 && per_image > 0, "bad argument");
    ig_reduce(grads, x, baseline, n_steps, per_image, out, (cudaStream_t)stream);
    SYNT_CATCH
}

int synt_intervene_blend(const float* x, const float* mask, const float* aux, int type, float noise_std, int B, int C,
                         int H, int W, float* out, void* stream) {
    return synt_intervene_blend_ex(x, mask, aux, type, noise_std, 5, B, C, H, W, out, nullptr, stream);
}
int synt_intervene_blend_ex(const float* x, const float* mask, const float* aux, int type, float noise_std, int blur_kernel,
                            int B, int C, int H, int W, float* out, float* intervention_out, void* stream) {
    SYNT_TRY
    SYNT_CHECK(x && mask && out && B > 0, "bad argument");
    intervene_blend(x, mask, aux, type, noise_std, blur_kernel, B, C, H, W, out, intervention_out, (cudaStream_t)stream);
    SYNT_CATCH
}
int synt_patch_mask_apply(const float* x, const unsigned char* pm, int n_masks, int C, int H, int W, int patch, float* out,
                          void* stream) {
    SYNT_TRY
    SYNT_CHECK(x && pm && out && n_masks > 0 && patch > 0 && H % patch == 0 && W % patch == 0, "bad argument");
    patch_mask_apply(x, pm, n_masks, C, H, W, patch, out, (cudaStream_t)stream);
    SYNT_CATCH
}

}  // extern "C"
