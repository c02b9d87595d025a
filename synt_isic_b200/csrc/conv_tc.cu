// tcgen05 / TMEM / TMA implicit-GEMM convolution for sm_100a (bf16 x bf16 -> fp32).
//
// This is the dense-contraction kernel of the SYNT_ISIC hot path: every UNet2D 3x3 / 1x1
// convolution (diffusers ResnetBlock2D / Downsample2D / Upsample2D / attention projections,
// reached from core/generator/image_generator.py:400) and every BN-folded ResNet18
// convolution (xai/XAI.py:433-436) runs through it.
//
// GEMM view     D[m, n] = sum_k A[m, k] * Wt[n, k]
//   m  : 128 output pixels per CTA = one TN x TH x TW box of the NHWC activation tensor
//   n  : BN output channels (64 / 128 / 256), accumulator = 128 lanes x BN fp32 TMEM columns
//   k  : one iteration = one filter tap x 64 input channels.  The A tile of an iteration is
//        ONE 4-D TMA box {64 ch, TW, TH, TN} of the activation tensor shifted by the tap
//        offset -- TMA's out-of-bounds zero fill IS the convolution padding, so no im2col
//        buffer exists anywhere.  Stride-2 convolutions use four "parity" tensor maps
//        (even/odd rows x even/odd columns of the input), 1x1 shortcut convolutions are
//        extra K iterations accumulated into the same TMEM tile (residual fusion).
//   smem: per stage A = 128 rows x 128 B and B = BN rows x 128 B, both K-major SWIZZLE_128B
//         exactly as the TMA writes them and as the UMMA shared-memory descriptor reads them.
//
// Warp roles (192 threads): warp 0 = TMA producer (one lane), warp 1 = TMEM allocator + MMA
// issuer (one lane), warps 2..5 = epilogue (TMEM -> registers -> bias/temb/residual/ReLU ->
// bf16 NHWC global stores).  Stages hand over through mbarriers; tcgen05.commit releases a
// stage back to the producer and signals the epilogue.
#include "kernels.cuh"
#include "ptx.cuh"
#include <cuda.h>
#include <mutex>

namespace synt {

using namespace ptx;

struct ConvTcMaps {
    CUtensorMap a[4];      // main input; [py*2+px] parity maps when stride == 2, a[0] otherwise
    CUtensorMap sc[2];     // 1x1 shortcut sources
    CUtensorMap b;         // weights [Cout][Ktot], K-major
};

struct ConvTcParams {
    int n_main_iters, cin_chunks, KW, pad, stride;
    int sc0_chunks, sc1_chunks;
    int tiles_x, tiles_y, TW, TH, TN;
    int B, Ho, Wo, Cout;
    const float* bias;
    const float* bias2;
    const bf16* residual;
    bf16* out;
    int relu;
};

constexpr int TC_THREADS = 192;
constexpr int TC_A_BYTES = 128 * 128;            // 128 pixels x 64 bf16

template <int BN, int STAGES>
struct TcSmem {
    static constexpr int B_BYTES = BN * 128;
    static constexpr int STAGE_BYTES = TC_A_BYTES + B_BYTES;
    static constexpr int BAR_OFFSET = STAGES * STAGE_BYTES;
    static constexpr int TOTAL = BAR_OFFSET + 256 + 1024;   // barriers + alignment slack
};

template <int BN, int STAGES>
__global__ void __launch_bounds__(TC_THREADS) conv_tc_kernel(const __grid_constant__ ConvTcMaps maps,
                                                             const __grid_constant__ ConvTcParams p) {
    using L = TcSmem<BN, STAGES>;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    uint8_t* smem = smem_raw + (((raw + 1023u) & ~1023u) - raw);
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::BAR_OFFSET);
    uint64_t* empty_bar = full_bar + STAGES;
    uint64_t* accum_bar = empty_bar + STAGES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accum_bar + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int total_iters = p.n_main_iters + p.sc0_chunks + p.sc1_chunks;

    // tile coordinates
    const int tiles_per_grp = p.tiles_x * p.tiles_y;
    const int tn = blockIdx.x / tiles_per_grp;
    const int trem = blockIdx.x - tn * tiles_per_grp;
    const int n0 = tn * p.TN, y0 = (trem / p.tiles_x) * p.TH, x0 = (trem % p.tiles_x) * p.TW;
    const int nt0 = blockIdx.y * BN;

    pdl_launch_dependents();
    if (warp == 0 && lane == 0) {
        prefetch_tmap(&maps.a[0]);
        prefetch_tmap(&maps.b);
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        mbar_init(accum_bar, 1);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc<(BN < 32 ? 32 : BN)>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_acc = *tmem_slot;
    pdl_wait();                                              // prologue done; the inputs come from the previous kernel

    if (warp == 0) {
        if (elect_one()) {
            // ===================== TMA producer =====================
            int stage = 0; uint32_t phase = 0;
            for (int it = 0; it < total_iters; ++it) {
                mbar_wait(&empty_bar[stage], phase ^ 1u);
                uint8_t* sa = smem + stage * L::STAGE_BYTES;
                uint8_t* sb = sa + TC_A_BYTES;
                mbar_arrive_expect_tx(&full_bar[stage], L::STAGE_BYTES);
                if (it < p.n_main_iters) {
                    const int tap = it / p.cin_chunks, chunk = it - tap * p.cin_chunks;
                    const int dy = tap / p.KW - p.pad, dx = tap % p.KW - p.pad;
                    if (p.stride == 1) {
                        tma_load_4d(sa, &maps.a[0], &full_bar[stage], chunk * 64, x0 + dx, y0 + dy, n0);
                    } else {
                        const int py = dy & 1, px = dx & 1;
                        tma_load_4d(sa, &maps.a[py * 2 + px], &full_bar[stage], chunk * 64, x0 + ((dx - px) >> 1),
                                    y0 + ((dy - py) >> 1), n0);
                    }
                } else {
                    int j = it - p.n_main_iters;
                    if (j < p.sc0_chunks) tma_load_4d(sa, &maps.sc[0], &full_bar[stage], j * 64, x0, y0, n0);
                    else tma_load_4d(sa, &maps.sc[1], &full_bar[stage], (j - p.sc0_chunks) * 64, x0, y0, n0);
                }
                tma_load_2d(sb, &maps.b, &full_bar[stage], it * 64, nt0);
                if (++stage == STAGES) { stage = 0; phase ^= 1u; }
            }
        }
    } else if (warp == 1) {
        if (elect_one()) {
            // ===================== MMA issuer =====================
            constexpr uint32_t idesc = make_idesc_bf16(128, BN);
            int stage = 0; uint32_t phase = 0;
            for (int it = 0; it < total_iters; ++it) {
                mbar_wait(&full_bar[stage], phase);
                tc_fence_after();
                const uint32_t sa = smem_u32(smem + stage * L::STAGE_BYTES);
                const uint64_t da = make_smem_desc_sw128(sa);
                const uint64_t db = make_smem_desc_sw128(sa + TC_A_BYTES);
#pragma unroll
                for (int k = 0; k < 4; ++k)     // 4 x UMMA_K(16) = 64 channels; +32 B = +2 in addr>>4 units
                    umma_bf16(tmem_acc, da + 2 * k, db + 2 * k, idesc, (it | k) != 0 ? 1u : 0u);
                umma_commit(&empty_bar[stage]);          // frees the smem stage when the MMAs retire
                if (++stage == STAGES) { stage = 0; phase ^= 1u; }
            }
            umma_commit(accum_bar);                      // accumulator complete -> epilogue
        }
    } else {
        // ===================== epilogue (warps 2..5) =====================
        const int q = warp & 3;                          // TMEM lane quarter this warp may access
        const int r = q * 32 + lane;                     // accumulator row == tile-local pixel
        const int ltn = r / (p.TH * p.TW), lrem = r % (p.TH * p.TW);
        const int bn = n0 + ltn, oy = y0 + lrem / p.TW, ox = x0 + lrem % p.TW;
        const bool valid = bn < p.B && oy < p.Ho && ox < p.Wo;
        const size_t pix = ((size_t)bn * p.Ho + oy) * p.Wo + ox;
        mbar_wait(accum_bar, 0);
        tc_fence_after();
#pragma unroll 1
        for (int c0 = 0; c0 < BN; c0 += 32) {
            uint32_t v[32];
            tmem_ld_32x32b_x32(tmem_acc + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, v);
            tmem_ld_wait();
            if (valid) {
                const int n = nt0 + c0;
                const size_t off = pix * p.Cout + n;
#pragma unroll
                for (int g = 0; g < 4; ++g) {            // 4 groups of 8 channels = 16 B of bf16 each
                    float f[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j) f[j] = __uint_as_float(v[g * 8 + j]) + __ldg(p.bias + n + g * 8 + j);
                    if (p.bias2) {
#pragma unroll
                        for (int j = 0; j < 8; ++j) f[j] += __ldg(p.bias2 + n + g * 8 + j);
                    }
                    if (p.residual) {
                        float rr[8];
                        load8<bf16>(p.residual + off + g * 8, rr);
#pragma unroll
                        for (int j = 0; j < 8; ++j) f[j] += rr[j];
                    }
                    if (p.relu) {
#pragma unroll
                        for (int j = 0; j < 8; ++j) f[j] = fmaxf(f[j], 0.f);
                    }
                    store8<bf16>(p.out + off + g * 8, f);
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc<(BN < 32 ? 32 : BN)>(tmem_acc);
}

// ------------------------------------------------------------------ host side ---------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
    static EncodeTiledFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* sym = nullptr;
        cudaDriverEntryPointQueryResult qres;
        cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres);
        if (e == cudaSuccess && qres == cudaDriverEntryPointSuccess) fn = reinterpret_cast<EncodeTiledFn>(sym);
    });
    if (!fn) throw Error(-4, "cuTensorMapEncodeTiled not available from the driver");
    return fn;
}

// NHWC bf16 activation [N, H, W, C] viewed with pixel strides (sy, sx) from a base offset:
// dims (C, W/sx, H/sy, N), box (64, TW, TH, TN), SWIZZLE_128B, OOB -> zero.
static void make_act_map(CUtensorMap* m, const bf16* base, int N, int H, int W, int C, int sy, int sx, int oy, int ox,
                         int TW, int TH, int TN) {
    const bf16* p = base + ((size_t)oy * W + ox) * C;
    cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)((W - ox + sx - 1) / sx), (cuuint64_t)((H - oy + sy - 1) / sy),
                          (cuuint64_t)N};
    cuuint64_t strides[3] = {(cuuint64_t)sx * C * 2, (cuuint64_t)sy * W * C * 2, (cuuint64_t)H * W * C * 2};
    cuuint32_t box[4] = {64, (cuuint32_t)TW, (cuuint32_t)TH, (cuuint32_t)TN};
    cuuint32_t es[4] = {1, 1, 1, 1};
    CUresult r = get_encode()(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<bf16*>(p), dims, strides, box, es,
                              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                              CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) throw Error(-4, "cuTensorMapEncodeTiled(activation) failed: " + std::to_string((int)r));
}
static void make_weight_map(CUtensorMap* m, const bf16* w, int Cout, int Ktot, int BN) {
    cuuint64_t dims[2] = {(cuuint64_t)Ktot, (cuuint64_t)Cout};
    cuuint64_t strides[1] = {(cuuint64_t)Ktot * 2};
    cuuint32_t box[2] = {64, (cuuint32_t)BN};
    cuuint32_t es[2] = {1, 1};
    CUresult r = get_encode()(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<bf16*>(w), dims, strides, box, es,
                              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                              CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) throw Error(-4, "cuTensorMapEncodeTiled(weight) failed: " + std::to_string((int)r));
}

bool conv_tc_supported(const ConvArgs& a) {
    if (a.Cin1 || a.gn_mode) return false;                 // channel-concat / fused-GroupNorm inputs: conv_tc2 only
    if (a.Cin % 64 || a.sc0_C % 64 || a.sc1_C % 64 || a.Cout % 64) return false;
    if (a.stride != 1 && a.stride != 2) return false;
    if (a.stride == 2 && ((a.H & 1) || (a.W & 1))) return false;
    if (a.sc_stride != 1 && a.sc_stride != 2) return false;
    if (a.KH != a.KW) return false;
    return true;
}

// pick the 128-pixel box TN x TH x TW for an Ho x Wo output plane
static void pick_tile(int Ho, int Wo, int& TW, int& TH, int& TN) {
    if (Wo >= 16)      { TW = 16; TH = 8;  TN = 1; }
    else if (Wo > 8)   { TW = 16; TH = 8;  TN = 1; }     // 14x14: two 16x8 boxes per image
    else               { TW = 8;  TH = 8;  TN = 2; }     // 7x7 / 8x8: two images per tile
    (void)Ho;
}

template <int BN, int STAGES>
static void launch_tc(const ConvTcMaps& maps, const ConvTcParams& p, dim3 grid, cudaStream_t s) {
    using L = TcSmem<BN, STAGES>;
    ensure_dynamic_smem((const void*)(conv_tc_kernel<BN, STAGES>), L::TOTAL);
    launch_pdl<true>(conv_tc_kernel<BN, STAGES>, grid, dim3(TC_THREADS), L::TOTAL, s, maps, p);
    SYNT_LAUNCH_CHECK();
}

void conv_tc(const ConvArgs& a, cudaStream_t s) {
    SYNT_CHECK(conv_tc_supported(a), "conv_tc: unsupported shape");
    SYNT_CHECK(a.bias != nullptr, "conv_tc: bias required");
    ConvTcParams p{};
    p.cin_chunks = a.Cin / 64;
    p.n_main_iters = a.KH * a.KW * p.cin_chunks;
    p.KW = a.KW; p.pad = a.pad; p.stride = a.stride;
    p.sc0_chunks = a.sc0_C / 64; p.sc1_chunks = a.sc1_C / 64;
    pick_tile(a.Ho, a.Wo, p.TW, p.TH, p.TN);
    p.tiles_x = ceil_div(a.Wo, p.TW); p.tiles_y = ceil_div(a.Ho, p.TH);
    p.B = a.B; p.Ho = a.Ho; p.Wo = a.Wo; p.Cout = a.Cout;
    p.bias = a.bias; p.bias2 = a.bias2; p.residual = (const bf16*)a.residual; p.out = (bf16*)a.out; p.relu = a.relu;
    const int m_tiles = ceil_div(a.B, p.TN) * p.tiles_x * p.tiles_y;

    int BN = 64;
    const int cands[3] = {256, 128, 64};
    for (int c : cands) {
        if (a.Cout % c) continue;
        BN = c;
        if ((long long)m_tiles * (a.Cout / c) >= 148) break;      // largest tile that still fills the SMs
    }

    ConvTcMaps maps;
    memset(&maps, 0, sizeof(maps));
    const bf16* in = (const bf16*)a.in;
    if (a.stride == 1) {
        make_act_map(&maps.a[0], in, a.B, a.H, a.W, a.Cin, 1, 1, 0, 0, p.TW, p.TH, p.TN);
        for (int i = 1; i < 4; ++i) maps.a[i] = maps.a[0];
    } else {
        for (int py = 0; py < 2; ++py)
            for (int px = 0; px < 2; ++px)
                make_act_map(&maps.a[py * 2 + px], in, a.B, a.H, a.W, a.Cin, 2, 2, py, px, p.TW, p.TH, p.TN);
    }
    const int Hs = a.Ho * a.sc_stride, Ws = a.Wo * a.sc_stride;
    if (a.sc0_C) make_act_map(&maps.sc[0], (const bf16*)a.sc0, a.B, Hs, Ws, a.sc0_C, a.sc_stride, a.sc_stride, 0, 0, p.TW, p.TH, p.TN);
    else maps.sc[0] = maps.a[0];
    if (a.sc1_C) make_act_map(&maps.sc[1], (const bf16*)a.sc1, a.B, Hs, Ws, a.sc1_C, a.sc_stride, a.sc_stride, 0, 0, p.TW, p.TH, p.TN);
    else maps.sc[1] = maps.a[0];
    make_weight_map(&maps.b, (const bf16*)a.weight, a.Cout, a.ktot(), BN);

    dim3 grid(m_tiles, a.Cout / BN);
    // Two stages for BN = 256 / 128 (measured, 1000-frame ResNet18 forward: 5.25 -> 4.97 ms): the kernel is not persistent, and
    // with 96 KB / 64 KB of shared memory two / three CTAs share an SM, so one CTA's prologue (TMEM allocation, pipeline fill)
    // and epilogue overlap another's main loop; the stages in flight per SM stay the same.
    if (BN == 256)      launch_tc<256, 2>(maps, p, grid, s);
    else if (BN == 128) launch_tc<128, 2>(maps, p, grid, s);
    else                launch_tc<64, 4>(maps, p, grid, s);
}

}  // namespace synt
