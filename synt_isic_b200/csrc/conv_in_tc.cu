// UNet conv_in (Conv2d(3, 64, 3, padding=1) on x_t, diffusers UNet2DModel.conv_in reached from
// core/generator/image_generator.py:400) on the tcgen05 path, bf16 production mode: x fp32 NCHW [B,3,128,128] -> NHWC bf16
// [B,128,128,64] with the GroupNorm statistics of the output (per-channel sum / sum of squares per 8-row band) fused.
//
// The FMA version (conv_in3_tiled_kernel, elementwise.cu) needs 3.6 GFLOP of fp32 FMA per step at B = 64 and ran at 158 us =
// 14% of the HBM roofline of what is a 147 MB streaming operation.  Here the 3x3x3 window becomes UMMA operands without an
// im2col tile (the operand-view trick of stem_tc.cu): every patch pixel is ONE 16-byte shared-memory entry
//   [c0h c1h c2h c0l c1l c2l 0 0]      xh = bf16(x), xl = bf16(x - xh): x = xh + xl to 2^-17
// eight consecutive pixels are one 8 x 16 B core matrix of the un-swizzled K-major layout, so the A operand of M tile t and
// tap (dy, dx) is the descriptor {start = patch + (128 t + dy*PW + dx) * 16, SBO = 128}.  One K step (16 values) is that
// entry TWICE (LBO = 0: both K halves read the same core matrix) against the tap's weight tile
//   k 0..2 = wh[c]   k 3..5 = wh[c]   k 8..10 = wl[c]   (else 0)        w = wh + wl
// i.e. xh*wh + xl*wh + xh*wl: fp32-grade products on the bf16 tensor path (only xl*wl, 2^-16 relative, is dropped).
// Accumulator row r of tile t is the output pixel with patch-linear index 128 t + r (two wrap-around columns per patch row
// are computed and discarded).
//
// One work item = an 8-row band of one image (1024 output pixels, a 10 x 130 patch): nine M tiles x nine taps = 81 MMAs
// (M128 N64 K16), three tiles per round; epilogue per round: TMEM -> +bias -> bf16 -> staging tile in shared memory ->
// coalesced 16-byte stores + per-channel statistics of the stored bf16 values.  Two persistent CTAs per SM interleave their
// load / MMA / epilogue phases; the next item's patch is requested into registers before the MMA wait.
#include "kernels.cuh"
#include "ptx.cuh"
#include <cstring>

namespace synt {

using namespace ptx;

constexpr int CIT_THREADS = 384;                           // 12 warps: four per accumulator tile of a round
constexpr int CIT_ROWS = 8, CIT_W = 128, CIT_PW = CIT_W + 2, CIT_PH = CIT_ROWS + 2;
constexpr int CIT_TILES = (CIT_ROWS - 1) * CIT_PW + CIT_W > 8 * 128 ? 9 : 8;      // rows 0 .. 7*130+127 = 1037 -> 9 tiles
constexpr int CIT_ROUND = 3;                               // tiles per round (TMEM: 3 x 64 columns)
constexpr int CIT_PATCH_PIX = CIT_TILES * 128 + 2 * CIT_PW + 2 + 1;   // furthest entry an A view touches (+1) = 1415
constexpr int CIT_PATCH_BYTES = ((CIT_PATCH_PIX * 16 + 127) / 128) * 128;
constexpr int CIT_W_BYTES = 9 * 64 * 16 * 2;               // nine tap tiles of 64 x 16 bf16
constexpr int CIT_STAGE_PITCH = 144;                       // bytes per staged pixel (64 bf16 + pad)
constexpr int CIT_STAGE_BYTES = 128 * CIT_STAGE_PITCH;     // 18432 per tile
constexpr int CIT_OFF_PATCH = 0;
constexpr int CIT_OFF_W = CIT_PATCH_BYTES;
constexpr int CIT_OFF_STAGE = CIT_OFF_W + CIT_W_BYTES;
constexpr int CIT_OFF_BIAS = CIT_OFF_STAGE + CIT_ROUND * CIT_STAGE_BYTES;
constexpr int CIT_OFF_RED = CIT_OFF_BIAS + 256;            // [12 warps... 4 pixel quarters x 3 tiles][64] float2 statistics scratch
constexpr int CIT_OFF_ROWPIX = CIT_OFF_RED + 12 * 64 * 8;  // [3 tiles][128 rows] output pixel of a staged row (or -1)
constexpr int CIT_OFF_BAR = CIT_OFF_ROWPIX + CIT_ROUND * 128 * 4;
constexpr int CIT_SMEM = CIT_OFF_BAR + 64;
constexpr int CIT_IN_PER_THREAD = (3 * CIT_PH * CIT_PW + CIT_THREADS - 1) / CIT_THREADS;   // 11 raw input values per thread
static_assert(CIT_TILES == 9 && CIT_TILES % CIT_ROUND == 0, "tile rounds");
static_assert(2 * (CIT_SMEM + 1024) <= 227 * 1024, "two CTAs per SM");

__device__ __forceinline__ uint64_t cit_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {   // SWIZZLE_NONE, K-major
    uint64_t d = 0;
    d |= static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4);
    d |= static_cast<uint64_t>(lbo_bytes >> 4) << 16;
    d |= static_cast<uint64_t>(sbo_bytes >> 4) << 32;
    d |= static_cast<uint64_t>(1) << 46;
    return d;
}

__global__ void __launch_bounds__(CIT_THREADS, 2) conv_in_tc_kernel(const float* __restrict__ x /* [B,3,128,128] */,
                                                                    const uint4* __restrict__ w_taps, const float* __restrict__ bias,
                                                                    int n_items, bf16* __restrict__ out /* [B,128,128,64] */,
                                                                    float2* __restrict__ stats /* [B][16][64] or null */) {
    extern __shared__ __align__(128) uint8_t cit_smem[];
    uint8_t* patch = cit_smem + CIT_OFF_PATCH;
    uint8_t* stage = cit_smem + CIT_OFF_STAGE;
    float* bias_s = reinterpret_cast<float*>(cit_smem + CIT_OFF_BIAS);
    float2* red = reinterpret_cast<float2*>(cit_smem + CIT_OFF_RED);
    int* row_pix = reinterpret_cast<int*>(cit_smem + CIT_OFF_ROWPIX);
    uint64_t* mma_bar = reinterpret_cast<uint64_t*>(cit_smem + CIT_OFF_BAR);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(mma_bar + 1);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    pdl_launch_dependents();
    // ---- once per CTA: weights, bias, zeroed patch (the entries beyond the 10 x 130 pixels stay zero), TMEM, barrier
    for (int i = threadIdx.x; i < CIT_W_BYTES / 16; i += CIT_THREADS)
        reinterpret_cast<uint4*>(cit_smem + CIT_OFF_W)[i] = __ldg(w_taps + i);
    if (threadIdx.x < 64) bias_s[threadIdx.x] = __ldg(bias + threadIdx.x);
    for (int i = threadIdx.x; i < CIT_PATCH_BYTES / 16; i += CIT_THREADS)
        reinterpret_cast<uint4*>(patch)[i] = make_uint4(0u, 0u, 0u, 0u);
    if (threadIdx.x == 0) { mbar_init(mma_bar, 1); fence_barrier_init(); }
    if (warp == 1) tmem_alloc<256>(tmem_slot);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    uint32_t phase = 0;
    pdl_wait();                                              // x is written by the previous step's scheduler epilogue

    // raw input of an item (3 x 10 x 130 values, zero outside the image), CIT_IN_PER_THREAD per thread, requested one item ahead
    float pre[CIT_IN_PER_THREAD];
    auto prefetch = [&](int item) {
        const int pb = item >> 4, y0 = (item & 15) * CIT_ROWS - 1;
        const float* img = x + (size_t)pb * 3 * 128 * 128;
#pragma unroll
        for (int j = 0; j < CIT_IN_PER_THREAD; ++j) {
            const int i = threadIdx.x + j * CIT_THREADS;
            pre[j] = 0.f;
            if (i < 3 * CIT_PH * CIT_PW) {
                const int c = i / (CIT_PH * CIT_PW), rem = i - c * (CIT_PH * CIT_PW), r = rem / CIT_PW, col = rem - r * CIT_PW;
                const int iy = y0 + r, ix = col - 1;
                if (iy >= 0 && iy < 128 && ix >= 0 && ix < 128) pre[j] = __ldg(img + ((size_t)c * 128 + iy) * 128 + ix);
            }
        }
    };
    if ((int)blockIdx.x < n_items) prefetch(blockIdx.x);

    for (int it = blockIdx.x; it < n_items; it += gridDim.x) {
        const int b = it >> 4, band = it & 15, oy0 = band * CIT_ROWS;
        // ---- A: hi / lo split of the patch into the 16-byte entries (value slot = channel for xh, 3 + channel for xl)
#pragma unroll
        for (int j = 0; j < CIT_IN_PER_THREAD; ++j) {
            const int i = threadIdx.x + j * CIT_THREADS;
            if (i < 3 * CIT_PH * CIT_PW) {
                const int c = i / (CIT_PH * CIT_PW), pix = i - c * (CIT_PH * CIT_PW);
                const bf16 h = __float2bfloat16_rn(pre[j]);
                const bf16 l = __float2bfloat16_rn(pre[j] - __bfloat162float(h));
                unsigned short* e = reinterpret_cast<unsigned short*>(patch + pix * 16);
                e[c] = __bfloat16_as_ushort(h);
                e[3 + c] = __bfloat16_as_ushort(l);
            }
        }
        fence_proxy_async();                                              // generic-proxy writes -> UMMA (async proxy)
        __syncthreads();
        float acc[4] = {0.f, 0.f, 0.f, 0.f};                              // statistics of this thread's channel pair: (sum, sumsq) x 2
        const int st_tile = warp >> 2, st_q = warp & 3;                   // statistics: tile of the round, pixel quarter
#pragma unroll 1
        for (int rd = 0; rd < CIT_TILES / CIT_ROUND; ++rd) {
            // ---- B: three M tiles x nine taps, one elected thread
            if (warp == 0) {
                if (elect_one()) {
                    tc_fence_after();
                    constexpr uint32_t idesc = make_idesc_bf16(128, 64);
                    const uint32_t a0 = smem_u32(patch), w0 = smem_u32(cit_smem + CIT_OFF_W);
#pragma unroll 1
                    for (int tap = 0; tap < 9; ++tap) {
                        const uint64_t db = cit_desc(w0 + tap * 2048, 1024, 128);
                        const int shift = (tap / 3) * CIT_PW + (tap % 3);
#pragma unroll
                        for (int t = 0; t < CIT_ROUND; ++t)
                            umma_bf16(tmem + t * 64, cit_desc(a0 + ((rd * CIT_ROUND + t) * 128 + shift) * 16, 0, 128), db, idesc,
                                      tap != 0 ? 1u : 0u);
                    }
                    umma_commit(mma_bar);
                }
                __syncwarp();
            }
            if (rd == CIT_TILES / CIT_ROUND - 1 && it + (int)gridDim.x < n_items) prefetch(it + gridDim.x);   // in flight from here on
            mbar_wait(mma_bar, phase);
            phase ^= 1u;
            tc_fence_after();
            // ---- C: accumulators -> +bias -> bf16 -> staging tile [pixel][64]
            {
                const int t = warp >> 2, quarter = warp & 3, r = quarter * 32 + lane;
                uint8_t* dst = stage + t * CIT_STAGE_BYTES + r * CIT_STAGE_PITCH;
                // rows that are no output pixel (the two wrap-around columns of a patch row, the tail of the last tile) are
                // staged as zeros -- they drop out of the statistics -- and marked -1 in the row table of the copy-out
                const int i = (rd * CIT_ROUND + t) * 128 + r, orow = i / CIT_PW, ocol = i - orow * CIT_PW;
                const bool valid = orow < CIT_ROWS && ocol < CIT_W;
                row_pix[t * 128 + r] = valid ? orow * 128 + ocol : -1;
#pragma unroll
                for (int half = 0; half < 2; ++half) {
                    uint32_t v[32];
                    tmem_ld_32x32b_x32(tmem + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(t * 64 + half * 32), v);
                    tmem_ld_wait();
#pragma unroll
                    for (int g = 0; g < 4; ++g) {
                        uint4 pk;
                        __nv_bfloat162* h2 = reinterpret_cast<__nv_bfloat162*>(&pk);
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const int ch = half * 32 + g * 8 + 2 * j;
                            h2[j] = __floats2bfloat162_rn(__uint_as_float(v[g * 8 + 2 * j]) + bias_s[ch],
                                                          __uint_as_float(v[g * 8 + 2 * j + 1]) + bias_s[ch + 1]);
                        }
                        *reinterpret_cast<uint4*>(dst + half * 64 + g * 16) = valid ? pk : make_uint4(0u, 0u, 0u, 0u);
                    }
                }
            }
            tc_fence_before();
            __syncthreads();
            // ---- D: statistics of the staged bf16 values (valid pixels only) and coalesced stores
            {
                const uint8_t* src = stage + st_tile * CIT_STAGE_BYTES;
#pragma unroll 8
                for (int k = 0; k < 32; ++k) {
                    const uint32_t u = *reinterpret_cast<const uint32_t*>(src + (st_q * 32 + k) * CIT_STAGE_PITCH + lane * 4);
                    const float v0 = __uint_as_float(u << 16), v1 = __uint_as_float(u & 0xffff0000u);
                    acc[0] += v0; acc[1] = fmaf(v0, v0, acc[1]); acc[2] += v1; acc[3] = fmaf(v1, v1, acc[3]);
                }
                const int tt = threadIdx.x & 127;                         // copy-out: 4 warps per tile, 8 chunks of 16 B per pixel
                bf16* band_out = out + ((size_t)b * 128 + oy0) * 128 * 64;
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    const int q = k * 128 + tt, r = q >> 3, chunk = q & 7;
                    const int pix = row_pix[st_tile * 128 + r];
                    if (pix >= 0)
                        *reinterpret_cast<uint4*>(band_out + (size_t)pix * 64 + chunk * 8) =
                            *reinterpret_cast<const uint4*>(src + r * CIT_STAGE_PITCH + chunk * 16);
                }
            }
            __syncthreads();                                             // staging tiles are rewritten by the next round
        }
        // ---- band statistics: 12 partial rows (3 tiles x 4 pixel quarters) per channel pair, summed in a fixed order
        if (stats) {
            red[warp * 64 + 2 * lane] = make_float2(acc[0], acc[1]);
            red[warp * 64 + 2 * lane + 1] = make_float2(acc[2], acc[3]);
            __syncthreads();
            if (threadIdx.x < 64) {
                float s0 = 0.f, s1 = 0.f;
#pragma unroll
                for (int w = 0; w < 12; ++w) { const float2 p = red[w * 64 + threadIdx.x]; s0 += p.x; s1 += p.y; }
                stats[((size_t)b * 16 + band) * 64 + threadIdx.x] = make_float2(s0, s1);
            }
            __syncthreads();
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc<256>(tmem);
}

// Host side: fp32 conv_in weights w[27][64] (k = tap*3 + c, the ConvInW packing) -> nine tap tiles in un-swizzled canonical
// layout: byte offset of (tap, n, k) = tap*2048 + (k/8)*1024 + (n/8)*128 + (n%8)*16 + (k%8)*2; k 0..2 and 3..5 = bf16(w),
// k 8..10 = bf16(w - bf16(w)).
static uint16_t cit_f2bf(float f) { uint32_t u; memcpy(&u, &f, 4); return (uint16_t)((u + 0x7FFFu + ((u >> 16) & 1u)) >> 16); }
static float cit_bf2f(uint16_t h) { uint32_t u = (uint32_t)h << 16; float f; memcpy(&f, &u, 4); return f; }
void conv_in_tc_pack_weights(const ConvInW& w, uint16_t* out /* CIT_W_BYTES / 2 */) {
    for (int i = 0; i < CIT_W_BYTES / 2; ++i) out[i] = 0;
    for (int tap = 0; tap < 9; ++tap)
        for (int n = 0; n < 64; ++n)
            for (int c = 0; c < 3; ++c) {
                const float f = w.w[tap * 3 + c][n];
                const uint16_t hi = cit_f2bf(f), lo = cit_f2bf(f - cit_bf2f(hi));
                auto at = [&](int k) { return ((size_t)tap * 2048 + (k / 8) * 1024 + (n / 8) * 128 + (n % 8) * 16 + (k % 8) * 2) / 2; };
                out[at(c)] = hi; out[at(3 + c)] = hi; out[at(8 + c)] = lo;
            }
}
int conv_in_tc_weight_bytes() { return CIT_W_BYTES; }

void conv_in_tc(const float* x_nchw, int B, const void* w_taps, const float* bias, void* out, float2* stats, cudaStream_t s) {
    ensure_dynamic_smem((const void*)(conv_in_tc_kernel), CIT_SMEM);
    static const int num_sms = [] { int dev = 0, n = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev); return n; }();
    const int n_items = B * 16;
    const int grid = n_items < 2 * num_sms ? n_items : 2 * num_sms;
    launch_pdl(conv_in_tc_kernel, dim3(grid), dim3(CIT_THREADS), CIT_SMEM, s, x_nchw, (const uint4*)w_taps, bias, n_items, (bf16*)out,
               stats);
}

}  // namespace synt
