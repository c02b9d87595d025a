// Classifier front end on the tcgen05 path (bf16 mode): preprocess -> conv 7x7 / stride 2 / pad 3 (3 -> 64, eval BN folded)
// + ReLU -> max-pool 3x3 / stride 2 / pad 1, ONE persistent kernel (reference: MelanomaClassifierAdaptive.preprocess_for_
// classifier + torchvision resnet18 conv1 / bn1 / relu / maxpool, xai/XAI.py:399-436).
//
// Why: the mma.sync version of this kernel (stem_fused_kernel, resnet.cu) sits at the legacy tensor path's ceiling on
// sm_100a (~144 TFLOP/s chip-wide; measured 1.07 ms per 512 images = 31% of the whole ResNet18 forward).
//
// How the 7x7 stride-2 window becomes UMMA operands WITHOUT an im2col tile.  The 224x224 preprocessed image is held in
// space-to-depth form: s2d pixel (Y, X) = the 2x2 block of image pixels (2Y+py, 2X+px), 12 values (py, px, c) padded to 16.
// The stride-2 7x7 convolution is then a stride-1 4x4 convolution over s2d pixels: output (oy, ox) reads s2d pixels
// (oy-2+ty, ox-2+tx), ty, tx = 0..3, with weight W[ky = 2ty+py-1][kx = 2tx+px-1] (zero where ky or kx falls outside 0..6).
// One tap = 16 values = exactly one K step of a bf16 tcgen05.mma.  The s2d patch of a tile is stored in shared memory as
// two planes of 16 bytes per pixel (values 0..7 / 8..15), pixels in row-major order: eight consecutive pixels of a plane
// are 128 contiguous bytes = one 8 x 16 B core matrix of the un-swizzled K-major canonical layout, so the A operand of M
// tile t and tap (ty, tx) is just the descriptor {start = plane0 + (128 t + ty*PW + tx) * 16, LBO = plane size, SBO = 128}:
// accumulator row r of tile t is the output pixel whose patch-linear index is 128 t + r (pixels whose column falls into
// the 3-pixel wrap-around margin of a patch row are computed and discarded).  B: per tap a 64 x 16 K-major tile (2 KB),
// all 16 taps resident in shared memory for the life of the CTA.
//
// One work item = 8x8 pooled pixels of one image = 17x17 conv pixels = a 20x20 s2d patch: three M tiles x 16 taps = 48
// MMAs (M128 N64 K16) into three 64-column accumulators; epilogue: TMEM -> +bias -> ReLU -> bf16 conv tile in shared memory
// -> 3x3/s2 max-pool -> global.  Two CTAs per SM interleave their sampling / MMA / epilogue phases.
//
// Measured (B200, 512 images): 0.57 ms against 1.07 ms for the mma.sync kernel.  Phase knock-outs (ms per 512 images): no
// sampling -0.18 (after staging the source window in shared memory and requesting it one item ahead; the exposed
// global-memory latency was another 0.15), one tap instead of 16 -0.14, no epilogue -0.13, no pooling -0.06.
#include "kernels.cuh"
#include "preprocess.cuh"
#include "ptx.cuh"

namespace synt {

using namespace ptx;

constexpr int STC_THREADS = 384;                           // 12 warps: four per accumulator tile in the epilogue
constexpr int STC_POOL = 8, STC_CONV = 2 * STC_POOL + 1;   // 8x8 pooled <- 17x17 conv pixels
constexpr int STC_PW = STC_CONV + 3;                       // 20 s2d pixels per patch row (4-tap window)
constexpr int STC_PATCH_PIX = STC_PW * STC_PW;             // 400 sampled s2d pixels
constexpr int STC_MT = 3;                                  // M tiles: rows 0 .. 16*20+16 = 336 of the patch-linear order
constexpr int STC_PLANE_PIX = STC_MT * 128 + 3 * STC_PW + 3 + 1;   // furthest pixel an A view touches (+1), 448
constexpr int STC_PLANE_BYTES = STC_PLANE_PIX * 16;        // 7168
constexpr int STC_W_BYTES = 16 * 64 * 16 * 2;              // 16 taps x (64 x 16) bf16 = 32768
constexpr int STC_TILE_PITCH = 144;                        // bytes per staged conv pixel (64 bf16 + pad, conflict-free)
constexpr int STC_OFF_PLANES = 0;
constexpr int STC_OFF_W = 2 * STC_PLANE_BYTES;             // 14336
constexpr int STC_OFF_TILE = STC_OFF_W + STC_W_BYTES;      // 47104
constexpr int STC_OFF_BIAS = STC_OFF_TILE + STC_CONV * STC_CONV * STC_TILE_PITCH;   // + 41616 = 88720
constexpr int STC_WIN_PER_THREAD = (3 * 25 * 25 + STC_THREADS - 1) / STC_THREADS;   // 5
constexpr int STC_WIN = 25, STC_WIN_PITCH = 27;            // source window: 40 rows of the 224 plane x 4/7 + the bilinear neighbour
constexpr int STC_OFF_WIN = STC_OFF_BIAS + 256;
constexpr int STC_OFF_LINES = STC_OFF_WIN + ((3 * STC_WIN * STC_WIN_PITCH * 4 + 127) / 128) * 128;   // 80 x float4
constexpr int STC_OFF_BAR = STC_OFF_LINES + 80 * 16;
constexpr int STC_SMEM = STC_OFF_BAR + 64;
static_assert(STC_PLANE_PIX == 448, "plane size");
static_assert(2 * (STC_SMEM + 1024) <= 227 * 1024, "two CTAs per SM");

// un-swizzled K-major operand: 8-row core matrices of 16-byte rows; LBO = distance between the two K halves of a K step,
// SBO = distance between consecutive 8-row groups (cute::UMMA::SmemDescriptor, LayoutType::SWIZZLE_NONE, version 1)
__device__ __forceinline__ uint64_t make_smem_desc_interleave(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4);
    d |= static_cast<uint64_t>(lbo_bytes >> 4) << 16;
    d |= static_cast<uint64_t>(sbo_bytes >> 4) << 32;
    d |= static_cast<uint64_t>(1) << 46;
    return d;
}

__global__ void __launch_bounds__(STC_THREADS, 2) stem_tc_kernel(const float* __restrict__ x /* [B,3,128,128] */,
                                                                 const uint4* __restrict__ w_taps /* STC_W_BYTES, canonical */,
                                                                 const float* __restrict__ bias, int n_items,
                                                                 bf16* __restrict__ out /* [B,56,56,64] */) {
    extern __shared__ __align__(128) uint8_t stc_smem[];
    uint8_t* planes = stc_smem + STC_OFF_PLANES;
    uint8_t* tile = stc_smem + STC_OFF_TILE;
    float* bias_s = reinterpret_cast<float*>(stc_smem + STC_OFF_BIAS);
    float* win = reinterpret_cast<float*>(stc_smem + STC_OFF_WIN);
    float4* lines = reinterpret_cast<float4*>(stc_smem + STC_OFF_LINES);
    uint64_t* mma_bar = reinterpret_cast<uint64_t*>(stc_smem + STC_OFF_BAR);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(mma_bar + 1);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    // ---- once per CTA: weights and bias into shared memory, margin pixels of the planes zeroed, TMEM, barrier
    for (int i = threadIdx.x; i < STC_W_BYTES / 16; i += STC_THREADS)
        reinterpret_cast<uint4*>(stc_smem + STC_OFF_W)[i] = __ldg(w_taps + i);
    if (threadIdx.x < 64) bias_s[threadIdx.x] = __ldg(bias + threadIdx.x);
    for (int i = threadIdx.x; i < STC_PLANE_PIX; i += STC_THREADS) {      // incl. the margin pixels and the pad values 12..15
        *reinterpret_cast<uint4*>(planes + i * 16) = make_uint4(0u, 0u, 0u, 0u);
        *reinterpret_cast<uint4*>(planes + STC_PLANE_BYTES + i * 16) = make_uint4(0u, 0u, 0u, 0u);
    }
    if (threadIdx.x == 0) { mbar_init(mma_bar, 1); fence_barrier_init(); }
    if (warp == 1) tmem_alloc<256>(tmem_slot);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    const float sc = 128.0f / 224.0f;
    uint32_t phase = 0;
    auto src_lo = [&](int g) { float f = ((float)max(g, 0) + 0.5f) * sc - 0.5f; return f < 0.f ? 0 : (int)f; };
    // the raw source window of an item, STC_WIN_PER_THREAD values per thread, requested one item ahead: the global-memory
    // latency (first touch of an image comes from HBM) is covered by the MMA / epilogue / pooling phases of the item before
    float pre[STC_WIN_PER_THREAD];
    auto prefetch = [&](int item) {
        const int pb = item / 49, ptl = item - pb * 49;
        const int pwy0 = src_lo(2 * (2 * ((ptl / 7) * STC_POOL) - 3)), pwx0 = src_lo(2 * (2 * ((ptl % 7) * STC_POOL) - 3));
        const float* pimg = x + (size_t)pb * 3 * 128 * 128;
#pragma unroll
        for (int j = 0; j < STC_WIN_PER_THREAD; ++j) {
            const int i = threadIdx.x + j * STC_THREADS;
            pre[j] = 0.f;
            if (i < 3 * STC_WIN * STC_WIN) {
                const int ch = i / (STC_WIN * STC_WIN), rem = i - ch * (STC_WIN * STC_WIN), r = rem / STC_WIN, c = rem - r * STC_WIN;
                pre[j] = __ldg(pimg + ((size_t)ch * 128 + min(pwy0 + r, 127)) * 128 + min(pwx0 + c, 127));
            }
        }
    };
    if ((int)blockIdx.x < n_items) prefetch(blockIdx.x);

    for (int it = blockIdx.x; it < n_items; it += gridDim.x) {
        const int b = it / 49, tl = it - b * 49;
        const int py0 = (tl / 7) * STC_POOL, px0 = (tl % 7) * STC_POOL;
        const int cy0 = 2 * py0 - 1, cx0 = 2 * px0 - 1;                   // conv-plane origin of the tile
        const int Y0 = cy0 - 2, X0 = cx0 - 2;                             // s2d origin of the patch
        // ---- A1: the source window of the patch (<= 25 x 25 pixels x 3 channels of the 128 x 128 image; fetched into registers
        //      one item ahead, see `prefetch`), clamp((x+1)/2, 0, 1) applied once per source pixel, staged in shared memory
        const int gy0 = 2 * Y0, gx0 = 2 * X0;                              // 224-plane origin of the patch (40 x 40 pixels)
        const int wy0 = src_lo(gy0), wx0 = src_lo(gx0);
#pragma unroll
        for (int j = 0; j < STC_WIN_PER_THREAD; ++j) {
            const int i = threadIdx.x + j * STC_THREADS;
            if (i < 3 * STC_WIN * STC_WIN) {
                const int ch = i / (STC_WIN * STC_WIN), rem = i - ch * (STC_WIN * STC_WIN), r = rem / STC_WIN, c = rem - r * STC_WIN;
                const float v = (pre[j] + 1.0f) / 2.0f;
                win[(ch * STC_WIN + r) * STC_WIN_PITCH + c] = fminf(fmaxf(v, 0.f), 1.f);
            }
        }
        // per patch row / column of the 224 plane (40 each): window offsets of the two source lines and the interpolation weight;
        // .w < 0 marks a line outside the image (the conv's zero padding)
        if (threadIdx.x < 80) {
            const int k = threadIdx.x, h = k >= 40, g = (h ? gx0 : gy0) + (h ? k - 40 : k);
            float f = ((float)g + 0.5f) * sc - 0.5f; if (f < 0.f) f = 0.f;
            const int a0 = (int)f, o0 = a0 - (h ? wx0 : wy0), o1 = o0 + (a0 < 127 ? 1 : 0);
            const bool ok = g >= 0 && g < 224;
            lines[k] = make_float4(__int_as_float(ok ? o0 : 0), __int_as_float(ok ? o1 : 0), f - (float)a0, ok ? 1.f : -1.f);
        }
        __syncthreads();
        // ---- A2: bilinear 128 -> 224 + ImageNet normalise from the staged window (the arithmetic of preprocess_pixel, with the
        //      division by the channel std as a multiplication).  One thread-pass = one pixel of the 224 plane = 3 values;
        //      s2d pixel (Y, X) keeps its 12 values (py, px, c) in two 16-byte plane entries (values 0..7 | 8..11, pad)
        {
            const float mean[3] = {0.485f, 0.456f, 0.406f};
            const float istd[3] = {1.0f / 0.229f, 1.0f / 0.224f, 1.0f / 0.225f};
#pragma unroll
            for (int pass = 0; pass < (4 * STC_PATCH_PIX + STC_THREADS - 1) / STC_THREADS; ++pass) {
                const int sidx = pass * STC_THREADS + threadIdx.x;          // sample index: pixel (gy, gx) of the 40 x 40 patch
                if (sidx < 4 * STC_PATCH_PIX) {
                    const int gy = sidx / 40, gx = sidx - gy * 40;
                    const float4 ry = lines[gy], rx = lines[40 + gx];
                    const int yo0 = __float_as_int(ry.x) * STC_WIN_PITCH, yo1 = __float_as_int(ry.y) * STC_WIN_PITCH;
                    const int xo0 = __float_as_int(rx.x), xo1 = __float_as_int(rx.y);
                    const bool ok = ry.w > 0.f && rx.w > 0.f;
                    const int pix = (gy >> 1) * STC_PW + (gx >> 1), q0 = ((gy & 1) * 2 + (gx & 1)) * 3;
#pragma unroll
                    for (int ch = 0; ch < 3; ++ch) {
                        const float* pl = win + ch * STC_WIN * STC_WIN_PITCH;
                        const float top = pl[yo0 + xo0] * (1.f - rx.z) + pl[yo0 + xo1] * rx.z;
                        const float bot = pl[yo1 + xo0] * (1.f - rx.z) + pl[yo1 + xo1] * rx.z;
                        const float o = ok ? ((top * (1.f - ry.z) + bot * ry.z) - mean[ch]) * istd[ch] : 0.f;
                        const int q = q0 + ch;
                        *reinterpret_cast<unsigned short*>(planes + (q >> 3) * STC_PLANE_BYTES + pix * 16 + (q & 7) * 2) =
                            __bfloat16_as_ushort(__float2bfloat16_rn(o));
                    }
                }
            }
        }
        fence_proxy_async();                                              // generic-proxy writes -> UMMA (async proxy)
        __syncthreads();
        // ---- B: 3 M tiles x 16 taps, one elected thread
        if (warp == 0) {
            if (elect_one()) {
                tc_fence_after();
                constexpr uint32_t idesc = make_idesc_bf16(128, 64);
                const uint32_t a0 = smem_u32(planes), w0 = smem_u32(stc_smem + STC_OFF_W);
#pragma unroll 1
                for (int tap = 0; tap < 16; ++tap) {
                    const uint64_t db = make_smem_desc_interleave(w0 + tap * 2048, 1024, 128);
                    const int shift = (tap >> 2) * STC_PW + (tap & 3);
#pragma unroll
                    for (int t = 0; t < STC_MT; ++t)
                        umma_bf16(tmem + t * 64, make_smem_desc_interleave(a0 + (t * 128 + shift) * 16, STC_PLANE_BYTES, 128), db,
                                  idesc, tap != 0 ? 1u : 0u);
                }
                umma_commit(mma_bar);
            }
            __syncwarp();
        }
        if (it + (int)gridDim.x < n_items) prefetch(it + gridDim.x);     // in flight during the MMAs, the epilogue and the pooling
        mbar_wait(mma_bar, phase);
        phase ^= 1u;
        tc_fence_after();
        // ---- C: accumulators -> +bias -> ReLU -> bf16 conv tile (pixels outside the 112x112 conv plane become 0 = the
        //      max-pool's padding: every window holds an in-plane value and all values are >= 0 after the ReLU)
        {
            const int t = warp >> 2, quarter = warp & 3;
            const int i = t * 128 + quarter * 32 + lane;                  // patch-linear index of this accumulator row
            const int oy = i / STC_PW, ox = i - oy * STC_PW;
            const bool keep = oy < STC_CONV && ox < STC_CONV;
            const int cy = cy0 + oy, cx = cx0 + ox;
            const bool in = cy >= 0 && cy < 112 && cx >= 0 && cx < 112;
            uint8_t* dst = tile + (oy * STC_CONV + ox) * STC_TILE_PITCH;
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                uint32_t v[32];
                tmem_ld_32x32b_x32(tmem + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(t * 64 + half * 32), v);
                tmem_ld_wait();
                if (keep) {
#pragma unroll
                    for (int g = 0; g < 4; ++g) {
                        uint4 pk;
                        __nv_bfloat162* h2 = reinterpret_cast<__nv_bfloat162*>(&pk);
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const int ch = half * 32 + g * 8 + 2 * j;
                            const float a = in ? fmaxf(__uint_as_float(v[g * 8 + 2 * j]) + bias_s[ch], 0.f) : 0.f;
                            const float bq = in ? fmaxf(__uint_as_float(v[g * 8 + 2 * j + 1]) + bias_s[ch + 1], 0.f) : 0.f;
                            h2[j] = __floats2bfloat162_rn(a, bq);
                        }
                        *reinterpret_cast<uint4*>(dst + half * 64 + g * 16) = pk;
                    }
                }
            }
        }
        tc_fence_before();
        __syncthreads();
        // ---- D: max-pool 3x3 / stride 2 / pad 1: pooled (py, px) <- conv tile rows 2py..2py+2, cols 2px..2px+2 (tile coords)
        for (int i = threadIdx.x; i < STC_POOL * STC_POOL * 8; i += STC_THREADS) {
            const int pp = i >> 3, ch = i & 7, py = pp / STC_POOL, px = pp % STC_POOL;
            __nv_bfloat162 mx[4];
#pragma unroll
            for (int dy = 0; dy < 3; ++dy)
#pragma unroll
                for (int dx = 0; dx < 3; ++dx) {
                    const uint4 v = *reinterpret_cast<const uint4*>(tile + ((2 * py + dy) * STC_CONV + 2 * px + dx) * STC_TILE_PITCH + ch * 16);
                    const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&v);
#pragma unroll
                    for (int j = 0; j < 4; ++j) mx[j] = (dy == 0 && dx == 0) ? h2[j] : __hmax2(mx[j], h2[j]);
                }
            *reinterpret_cast<uint4*>(out + (((size_t)b * 56 + py0 + py) * 56 + px0 + px) * 64 + ch * 8) = *reinterpret_cast<const uint4*>(mx);
        }
        // the next item's sampling overwrites the planes (their MMAs are complete) and, after its barrier, the conv tile
        // (every thread has left this loop by then)
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc<256>(tmem);
}

// Host side: BN-folded stem weights [64][ky*24 + kx*3 + c] (bf16 bits, the K-major packing of resnet.cu) -> the 16 tap tiles
// in un-swizzled canonical layout: byte offset of (tap, n, k) = tap*2048 + (k/8)*1024 + (n/8)*128 + (n%8)*16 + (k%8)*2 with
// k = (py*2 + px)*3 + c (k >= 12 zero), ky = 2*ty + py - 1, kx = 2*tx + px - 1.
void stem_tc_pack_weights(const uint16_t* w_k192 /* [64][192] */, uint16_t* out /* STC_W_BYTES / 2 */) {
    for (int i = 0; i < STC_W_BYTES / 2; ++i) out[i] = 0;
    for (int tap = 0; tap < 16; ++tap) {
        const int ty = tap >> 2, tx = tap & 3;
        for (int n = 0; n < 64; ++n)
            for (int k = 0; k < 12; ++k) {
                const int py = k / 6, px = (k / 3) % 2, c = k % 3;
                const int ky = 2 * ty + py - 1, kx = 2 * tx + px - 1;
                if (ky < 0 || ky > 6 || kx < 0 || kx > 6) continue;
                const size_t off = (size_t)tap * 2048 + (k / 8) * 1024 + (n / 8) * 128 + (n % 8) * 16 + (k % 8) * 2;
                out[off / 2] = w_k192[(size_t)n * 192 + ky * 24 + kx * 3 + c];
            }
    }
}
int stem_tc_weight_bytes() { return STC_W_BYTES; }

void stem_tc(const float* x_nchw, int B, const void* w_taps, const float* bias, void* out, cudaStream_t s) {
    ensure_dynamic_smem((const void*)(stem_tc_kernel), STC_SMEM);
    static const int num_sms = [] { int dev = 0, n = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev); return n; }();
    const int n_items = B * 49;
    const int grid = n_items < 2 * num_sms ? n_items : 2 * num_sms;
    stem_tc_kernel<<<grid, STC_THREADS, STC_SMEM, s>>>(x_nchw, (const uint4*)w_taps, bias, n_items, (bf16*)out);
    SYNT_LAUNCH_CHECK();
}

}  // namespace synt
