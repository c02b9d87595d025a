// 3x3 stride-1 implicit-GEMM convolution with ONE halo-tile TMA load per 64-channel chunk.
//
// conv_tc.cu loads the 128-pixel A tile nine times per input-channel chunk (once per filter tap,
// shifted by the tap offset).  Here the (16+2) x (8+2) pixel halo of the tile is loaded ONCE and
// the nine taps are nine shifted *views* of it: tile row m = (h, w) of tap (dy, dx) is halo pixel
// (h+dy, w+dx), i.e. the UMMA shared-memory descriptor starts (dy*pitch + dx)*128 bytes into the
// halo buffer and steps `pitch`*128 bytes between 8-row groups (the 8 pixels of one tile row).
// L2 -> shared-memory traffic for A drops from 9 x 16 KB to one halo (36 KB at pitch 16, 23 KB at
// pitch 10) per chunk, which is what bounds the Cout = 64 layers at 128x128.
//
// The shifted views are not 1024-byte aligned, so whether they are legal depends on how the
// tensor core applies the 128-byte swizzle (absolute address bits vs. row index + base_offset);
// `variant` selects the layout/descriptor flavour so that the hardware can be asked:
//   0: pitch 16 pixels (group stride 2048 B), base_offset field = 0
//   1: pitch 16 pixels,                      base_offset field = dx
//   2: pitch 10 pixels (group stride 1280 B), base_offset field = 0
//   3: pitch 10 pixels,                      base_offset field = (start_address >> 7) & 7
#include "kernels.cuh"
#include "ptx.cuh"
#include "tmap.cuh"

namespace synt {

using namespace ptx;

struct HaloMaps { CUtensorMap a; CUtensorMap b; };
struct HaloParams {
    int cin_chunks, tiles_x, tiles_y, B, H, W, Cout, pitch, variant;
    const float* bias; const float* bias2; const bf16* residual; bf16* out; int relu;
};

constexpr int HL_THREADS = 192;
constexpr int HL_BN = 64;
constexpr int HL_STAGES = 2;
constexpr int HL_A_BYTES = 18 * 16 * 128;                 // sized for the pitch-16 halo
constexpr int HL_B_BYTES = 9 * HL_BN * 128;
constexpr int HL_STAGE_BYTES = HL_A_BYTES + HL_B_BYTES;   // 110,592
constexpr int HL_SMEM = HL_STAGES * HL_STAGE_BYTES + 256 + 1024;

__device__ __forceinline__ uint64_t make_desc_sw128_ex(uint32_t addr, uint32_t sbo, uint32_t base_off) {
    uint64_t d = 0;
    d |= static_cast<uint64_t>((addr & 0x3FFFFu) >> 4);
    d |= static_cast<uint64_t>(1) << 16;
    d |= static_cast<uint64_t>(sbo >> 4) << 32;
    d |= static_cast<uint64_t>(1) << 46;
    d |= static_cast<uint64_t>(base_off & 7u) << 49;
    d |= static_cast<uint64_t>(2) << 61;
    return d;
}

__global__ void __launch_bounds__(HL_THREADS, 1) conv_halo_kernel(const __grid_constant__ HaloMaps maps,
                                                                  const __grid_constant__ HaloParams p) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    uint8_t* smem = smem_raw + (((raw + 1023u) & ~1023u) - raw);
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + HL_STAGES * HL_STAGE_BYTES);
    uint64_t* empty_bar = full_bar + HL_STAGES;
    uint64_t* accum_bar = empty_bar + HL_STAGES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accum_bar + 1);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    const int per_img = p.tiles_x * p.tiles_y;
    const int n0 = blockIdx.x / per_img, trem = blockIdx.x % per_img;
    const int y0 = (trem / p.tiles_x) * 16, x0 = (trem % p.tiles_x) * 8;
    const int nt0 = blockIdx.y * HL_BN;
    const int a_bytes = 18 * p.pitch * 128;

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&maps.a); prefetch_tmap(&maps.b);
        for (int s = 0; s < HL_STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        mbar_init(accum_bar, 1);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc<HL_BN>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_acc = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            int stage = 0; uint32_t phase = 0;
            for (int ch = 0; ch < p.cin_chunks; ++ch) {
                mbar_wait(&empty_bar[stage], phase ^ 1u);
                uint8_t* sa = smem + stage * HL_STAGE_BYTES;
                mbar_arrive_expect_tx(&full_bar[stage], a_bytes + HL_B_BYTES);
                tma_load_4d(sa, &maps.a, &full_bar[stage], ch * 64, x0 - 1, y0 - 1, n0);       // OOB -> 0 = conv padding
                for (int tap = 0; tap < 9; ++tap)
                    tma_load_2d(sa + HL_A_BYTES + tap * HL_BN * 128, &maps.b, &full_bar[stage],
                                (tap * p.cin_chunks + ch) * 64, nt0);
                if (++stage == HL_STAGES) { stage = 0; phase ^= 1u; }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            constexpr uint32_t idesc = make_idesc_bf16(128, HL_BN);
            const uint32_t sbo = p.pitch * 128;
            int stage = 0; uint32_t phase = 0;
            for (int ch = 0; ch < p.cin_chunks; ++ch) {
                mbar_wait(&full_bar[stage], phase);
                tc_fence_after();
                const uint32_t sa = smem_u32(smem + stage * HL_STAGE_BYTES);
                for (int tap = 0; tap < 9; ++tap) {
                    const int dy = tap / 3, dx = tap % 3;
                    const uint32_t a_addr = sa + (dy * p.pitch + dx) * 128;
                    uint32_t boff = 0;
                    if (p.variant == 1) boff = dx;
                    else if (p.variant == 3) boff = (a_addr >> 7) & 7u;
                    const uint64_t da = make_desc_sw128_ex(a_addr, sbo, boff);
                    const uint64_t db = make_smem_desc_sw128(sa + HL_A_BYTES + tap * HL_BN * 128);
#pragma unroll
                    for (int k = 0; k < 4; ++k) umma_bf16(tmem_acc, da + 2 * k, db + 2 * k, idesc, (ch | tap | k) != 0 ? 1u : 0u);
                }
                umma_commit(&empty_bar[stage]);
                if (++stage == HL_STAGES) { stage = 0; phase ^= 1u; }
            }
            umma_commit(accum_bar);
        }
    } else {
        const int q = warp & 3, r = q * 32 + lane;
        const int oy = y0 + r / 8, ox = x0 + r % 8;
        const bool valid = n0 < p.B && oy < p.H && ox < p.W;
        const size_t pix = ((size_t)n0 * p.H + oy) * p.W + ox;
        mbar_wait(accum_bar, 0);
        tc_fence_after();
#pragma unroll 1
        for (int c0 = 0; c0 < HL_BN; c0 += 32) {
            uint32_t v[32];
            tmem_ld_32x32b_x32(tmem_acc + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, v);
            tmem_ld_wait();
            if (valid) {
                const int n = nt0 + c0;
                const size_t off = pix * p.Cout + n;
#pragma unroll
                for (int g = 0; g < 4; ++g) {
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
    if (warp == 1) tmem_dealloc<HL_BN>(tmem_acc);
}

void conv_tc_halo(const ConvArgs& a, int variant, cudaStream_t s) {
    SYNT_CHECK(a.KH == 3 && a.KW == 3 && a.stride == 1 && a.pad == 1 && a.Cin % 64 == 0 && a.Cout % 64 == 0 &&
                   a.sc0_C == 0 && a.sc1_C == 0, "conv_tc_halo: 3x3 stride-1 only");
    HaloParams p{};
    p.cin_chunks = a.Cin / 64; p.tiles_x = ceil_div(a.W, 8); p.tiles_y = ceil_div(a.H, 16);
    p.B = a.B; p.H = a.H; p.W = a.W; p.Cout = a.Cout; p.variant = variant; p.pitch = variant < 2 ? 16 : 10;
    p.bias = a.bias; p.bias2 = a.bias2; p.residual = (const bf16*)a.residual; p.out = (bf16*)a.out; p.relu = a.relu;
    HaloMaps maps;
    {
        cuuint64_t dims[4] = {(cuuint64_t)a.Cin, (cuuint64_t)a.W, (cuuint64_t)a.H, (cuuint64_t)a.B};
        cuuint64_t strides[3] = {(cuuint64_t)a.Cin * 2, (cuuint64_t)a.W * a.Cin * 2, (cuuint64_t)a.H * a.W * a.Cin * 2};
        cuuint32_t box[4] = {64, (cuuint32_t)p.pitch, 18, 1};
        encode_bf16_sw128(&maps.a, a.in, 4, dims, strides, box, "halo activation");
    }
    {
        cuuint64_t dims[2] = {(cuuint64_t)a.ktot(), (cuuint64_t)a.Cout};
        cuuint64_t strides[1] = {(cuuint64_t)a.ktot() * 2};
        cuuint32_t box[2] = {64, HL_BN};
        encode_bf16_sw128(&maps.b, a.weight, 2, dims, strides, box, "halo weight");
    }
    ensure_dynamic_smem((const void*)(conv_halo_kernel), HL_SMEM);
    dim3 grid(a.B * p.tiles_x * p.tiles_y, a.Cout / HL_BN);
    conv_halo_kernel<<<grid, HL_THREADS, HL_SMEM, s>>>(maps, p);
    SYNT_LAUNCH_CHECK();
}

}  // namespace synt
