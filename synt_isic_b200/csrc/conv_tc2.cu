// Persistent tcgen05 implicit-GEMM kernel for the 3x3 stride-1 convolutions of the UNet
// (85% of the path's FLOPs; diffusers ResnetBlock2D.conv1/conv2 and Upsample2D.conv reached from
// core/generator/image_generator.py:400), including the fused 1x1 shortcut segments.
//
// Why a second kernel: conv_tc.cu (one CTA per 128-pixel tile, one TMA box per filter tap) moves
// 16 KB (A) + BN*128 B (B) from L2 per 64-deep K step, i.e. 96 B/clk/SM at full MMA rate for
// BN = 256 -- the measured ceiling is ~40 B/clk/SM, so it runs at ~40% of the tensor pipe.  Here
//   * A: ONE halo-tile TMA load per 64-channel chunk serves all nine taps of TWO vertically
//     adjacent 16x8-pixel tiles: the (34 x 10 pixel) halo is a SWIZZLE_128B smem tile whose nine
//     tap views are plain address offsets ((dy*10+dx)*128 B, 8-row group stride 1280 B) of the
//     UMMA descriptor -- measured legal on B200 (the 128B swizzle is a function of absolute
//     shared-memory address bits; tools/exp_halo.py).  A traffic drops 9 x 32 KB -> 43.5 KB.
//   * B: each weight tile (one tap x 64 channels x BN) feeds the MMAs of both M tiles.
//   * persistent CTAs (one per SM) with double-buffered TMEM accumulators: the epilogue of super-
//     tile i (TMEM -> registers -> bias/temb/residual -> bf16 -> swizzled smem -> TMA store)
//     overlaps the mainloop of super-tile i+1.
//   * GroupNorm(+SiLU) of the INPUT is applied in shared memory: four transform warps rewrite each
//     halo tile in place (y = act(x*scale[n,c] + shift[n,c]); out-of-image pixels stay 0 = the conv
//     padding of the activated tensor) between the TMA landing and the MMAs, once per input element
//     -- the normalised activation tensor is never materialised in HBM.
// L2 -> smem bytes per MMA clock: (43.5 KB + 9*BN*128 B) / (36*BN clk) = 42 B/clk for BN = 128.
//
// Warp roles (640 threads): warp 0 TMA producer of the activation tiles, warp 2 TMA producer of the weight tiles, warp 1 TMEM
// allocator + MMA issuer (one elected lane each), warps 4..7 / 8..11 two epilogue warpgroups (warpgroup e owns M tile e of every super-tile, its own
// staging tile and TMA stores; the residual tile is TMA-loaded INTO the staging tile and updated in place),
// warps 12..19 input transform.  The roles are latency-bound (one warp per SM sub-partition each would
// leave the tensor pipe waiting), hence two warps per sub-partition for epilogue and transform.
// Rings: A halo slots (2), B weight-tile slots (NB), TMEM accumulator buffers (2).
#include "kernels.cuh"
#include "ptx.cuh"
#include <type_traits>
#include <vector>
#include "tmap.cuh"

namespace synt {

using namespace ptx;

constexpr int V2_THREADS = 640;
constexpr int V2_EPI_BASE = 128;                           // epilogue warps 4..11 (two warpgroups, one per M tile)
constexpr int V2_XF_BASE = 384;                            // transform warps 12..19
constexpr int V2_XF_THREADS = 256;
constexpr int V2_MT = 2;                                   // M tiles (16x8 pixels each) per super-tile
constexpr int V2_A_STAGES = 2;

// RES = true: the whole weight matrix of the layer (<= 12 K blocks of 64 x BN) stays resident in shared
// memory for the life of the persistent CTA (Cout = 64 layers with Ktot <= 768: L2 traffic is A only).
// K1 = true: 1x1 projections (qkv, attention out-projection; BN = 128): no halo -- an A slot is exactly the 256 pixels of the
// super-tile (32 KB) -- and FOUR slots.  A chunk of a 1x1 conv feeds only 8 MMAs (~700 clk) while its load takes ~2500 clk, so
// the 2-deep ring of the 3x3 geometry ran these layers at the load latency (measured: the MMA issuer waits 35% of its time for
// a_ready and the qkv projection takes 59 us against ~20 us of HBM time); the staging tiles shrink to 64-column passes to pay
// for the two extra slots.
template <int BN, bool RES, bool K1 = false>
struct V2Smem {
    static constexpr int A_STAGES = K1 ? 4 : V2_A_STAGES;
    static constexpr int A_SLOT = K1 ? 32768 : 46080;      // 46080 = max(34*10, 2*18*10) * 128, already 1 KB aligned
    static constexpr int B_TILE = BN * 128;
    // BN = 256 (Cout = 256, large K): one UMMA of N = 256 per M tile -- per MMA the tensor core reads 4 KB of A and
    // BN*32 B of B from shared memory, and that operand traffic (measured ~90 B/clk) is what paces N <= 128 tiles
    // (98 clk per N = 128 UMMA against a 64 clk floor); N = 256 runs at its 128 clk floor.  The price: the two
    // accumulators of a super-tile fill all 512 TMEM columns (NBUF = 1, the epilogue does not overlap the next
    // mainloop) and the epilogue walks the tile in EC = 64-column passes so that the staging tiles stay small.
    static constexpr int EC = (BN == 256 || K1) ? 64 : BN; // columns per epilogue pass
    static constexpr int NPASS = BN / EC;
    static constexpr int NBUF = BN == 256 ? 1 : 2;         // TMEM accumulator buffers (each V2_MT x BN columns)
    static constexpr int STAGING = 128 * EC * 2;           // one pass of one M tile of bf16 output (one per epilogue warpgroup)
    static constexpr int NB = RES ? 12 : (BN == 256 ? 3 : (BN == 128 ? 4 : 12));   // BN = 64 streamed: 12 x 8 KB fit next to the A slots (b_full waits 9% with 8)
    static constexpr int OFF_B = A_STAGES * A_SLOT;
    static constexpr int OFF_STAGING = OFF_B + NB * B_TILE;
    static constexpr int OFF_BIAS = OFF_STAGING + V2_MT * STAGING;
    static constexpr int OFF_BAR = OFF_BIAS + V2_MT * BN * 4;
    static constexpr int TOTAL = OFF_BAR + 512 + 1024;
    static_assert(TOTAL <= 227 * 1024, "shared memory budget");
    static_assert(NBUF * V2_MT * BN <= 512, "TMEM columns");
};

// Role knock-outs (profiles/r01c_experiments.md): compiled only with -DSYNT_EXPERIMENTS (SYNT_EXPERIMENTS=1 python -m
// synt_isic_b200.build); the product kernel carries no run-time experiment tests.
#ifdef SYNT_EXPERIMENTS
#define V2_EXP(mask) ((p.exp_nob & (mask)) != 0)
// wait-time accounting of one thread per role: V2_TW(i, wait) adds the cycles spent in `wait` to record i of this CTA
#define V2_TW(i, stmt) do { if (p.tl) { const long long _t = clock64(); stmt; tlacc[i] += clock64() - _t; } else { stmt; } } while (0)
#define V2_TL_DECL long long tlacc[4] = {0, 0, 0, 0}; const long long tl_t0 = clock64()
#define V2_TL_FLUSH(base, n) do { if (p.tl) { for (int _i = 0; _i < (n); ++_i) p.tl[blockIdx.x * 16 + (base) + _i] = tlacc[_i]; \
        p.tl[blockIdx.x * 16 + (base) + (n)] = clock64() - tl_t0; } } while (0)
#else
#define V2_EXP(mask) (false)
#define V2_TW(i, stmt) do { stmt; } while (0)
#define V2_TL_DECL do { } while (0)
#define V2_TL_FLUSH(base, n) do { } while (0)
#endif

struct V2Maps { CUtensorMap a[4]; CUtensorMap b; CUtensorMap out[4]; CUtensorMap res; };   // a[i]: source of segment i; out[phase]
// A K segment = one source tensor: `chunks` 64-channel blocks x `taps` (9 = 3x3 window, 1 = centre tap);
// weight K block of (tap, chunk) = kb_base + tap*kb_stride + chunk; xform: 0 raw, 1 GroupNorm affine,
// 2 affine + SiLU with scale/shift rows gn_ss[n][ss_off + channel].
struct V2Seg { int chunks, taps, kb_base, kb_stride, xform, ss_off; };
struct V2Params {
    int n_work;              // super-tiles x N tiles
    int n_ntiles, tiles_x, supers_per_img, imgs_per_super, row_off;
    int nt_real;             // real N tiles; n_ntiles = nt_real * phases (phase = sub-pixel position of a fused 2x upsample)
    int n_seg; V2Seg seg[4];
    const float2* gn_ss; int gn_C;     // GroupNorm scale/shift [B][gn_C] of the (concatenated) main input
    int B, H, W, Cout;
    const float* bias; const float* bias2; int has_res; int relu;   // residual tile arrives through maps.res
    float2* stats; int stats_slots;   // optional fused GroupNorm partials [B][stats_slots][Cout]
    int chunk;                         // consecutive work items per CTA turn (divides the super-tiles per image)
    uint32_t mg_per_img, mg_ntiles, mg_nt_real, mg_tiles_x, mg_chunk;   // v2_div magic numbers of the decode's divisors
    long long* tl;                     // EXPERIMENTS: per-CTA wait-time records [grid][16] (SYNT_CONV_TL=1, tools/conv_timeline.py)
    int exp_nob;                       // EXPERIMENTS (SYNT_EXP_NOB bit mask, wrong results): 1 skip the weight-tile TMA loads after the
                                       // first ring fill, 2 skip the input transform, 4 skip the statistics pass, 8 TMA stores, 16 A-tile loads, 32 epilogue
                                       // body, 64 no producer/transform at all and no operand waits in the MMA issuer (pure MMA issue rate),
                                       // 128 transform copies without arithmetic, 256 transform without the store-back, 512 transform without SiLU
};

__device__ __forceinline__ void tma_store_4d(const CUtensorMap* m, const void* src, int c0, int c1, int c2, int c3) {
    asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
                 ::"l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
                 : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ float silu_tanh_v2(float x) {      // x*sigmoid(x) = h + h*tanh(h), h = x/2 (one MUFU op)
    const float h = 0.5f * x;
    float t;
    asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(h));
    return fmaf(h, t, h);
}
__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, 128;" ::: "memory"); }

// Work item w = ((image group) * n_ntiles + nt) * super_tiles_per_group + tile.  CTAs take CHUNKS of R
// consecutive items round-robin (chunk c -> CTA c % grid): neighbouring CTAs work on neighbouring tiles
// (L2/DRAM locality) while each chunk stays inside one (image, N tile), so GroupNorm partial sums are
// carried in registers across the chunk and written once -- slot = chunk index within the image.
struct V2Work { int n0, y0, x0, nt, grp, phase, ntr; };   // nt: virtual N tile (weights), ntr: real N tile (channels)
// Division by a launch-invariant divisor d: q = umulhi(n, 2^32 / d + 1), exact while n * d < 2^32 (work-item indices are
// < 2^20, divisors < 2^12).  Every thread of every role decodes its work items: with emulated integer division that was
// ~400 instructions per thread and item, i.e. ~2000 issue slots per sub-partition and item for the 20 warps of the CTA.
__device__ __forceinline__ int v2_div(int n, uint32_t magic, int d) { return d == 1 ? n : (int)__umulhi((uint32_t)n, magic); }
__device__ __forceinline__ V2Work v2_decode(const V2Params& p, int w) {
    V2Work o;
    const int per_img = p.tiles_x * p.supers_per_img;
    const int key = v2_div(w, p.mg_per_img, per_img), rem = w - key * per_img;
    o.grp = v2_div(key, p.mg_ntiles, p.n_ntiles);
    o.nt = key - o.grp * p.n_ntiles;
    o.phase = v2_div(o.nt, p.mg_nt_real, p.nt_real);
    o.ntr = o.nt - o.phase * p.nt_real;
    o.n0 = o.grp * p.imgs_per_super;
    const int ty = v2_div(rem, p.mg_tiles_x, p.tiles_x);
    o.y0 = ty * (p.imgs_per_super == 1 ? 32 : 0);
    o.x0 = (rem - ty * p.tiles_x) * 8;
    return o;
}
// iteration i of this CTA -> work item (or -1 when exhausted)
__device__ __forceinline__ int v2_item(const V2Params& p, int i) {
    const int turn = v2_div(i, p.mg_chunk, p.chunk);
    const int chunk = turn * (int)gridDim.x + (int)blockIdx.x;
    const int w = chunk * p.chunk + (i - turn * p.chunk);
    return w < p.n_work ? w : -1;
}

// PAIR: the two CTAs of a cluster (one TPC) work on two consecutive super-tiles of the same N tile and share every
// weight tile: each loads HALF of it (BN/2 rows) and the leader issues `tcgen05.mma.cta_group::2` (M = 256) for both,
// so per MMA an SM fetches 4 KB of A and only BN*16 B of B from shared memory.  The peer's MMA warp is a relay: it
// forwards "my A tile is transformed / my half of B has landed / my epilogue drained the accumulators" to the leader
// with remote mbarrier arrives; the leader's commits are multicast to both CTAs' empty / full barriers.
template <int BN, bool RES, bool PAIR = false, bool K1 = false>
__global__ void __launch_bounds__(V2_THREADS, 1) conv_tc2_kernel(const __grid_constant__ V2Maps maps,
                                                                 const __grid_constant__ V2Params p, bf16* __restrict__ out) {
    using L = V2Smem<BN, RES, K1>;
    constexpr int A_STAGES = L::A_STAGES;
    static_assert(!PAIR || (!RES && L::NBUF == 2 && !K1), "pair mode: streamed weights, double-buffered accumulators");
    const uint32_t rank = PAIR ? cluster_ctarank() : 0u;
    // iteration i of this CTA -> work item (or -1 when exhausted); pair mode: cluster c takes pair-items c, c + #clusters, ...
    auto v2_item = [&](const V2Params& pp, int i) -> int {
        if (!PAIR) return synt::v2_item(pp, i);
        const int w = 2 * (i * ((int)gridDim.x >> 1) + ((int)blockIdx.x >> 1)) + (int)rank;
        return w < pp.n_work ? w : -1;
    };
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    uint8_t* smem = smem_raw + (((raw + 1023u) & ~1023u) - raw);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L::OFF_BAR);
    uint64_t* a_full = bars;                        // [2]
    uint64_t* a_empty = a_full + A_STAGES;       // [2]
    uint64_t* a_ready = a_empty + A_STAGES;      // [2] halo tile transformed (or passed through)
    uint64_t* b_full = a_ready + A_STAGES;       // [NB]
    uint64_t* b_empty = b_full + L::NB;             // [NB]
    uint64_t* t_full = b_empty + L::NB;             // [2]
    uint64_t* t_empty = t_full + 2;                 // [2]
    uint64_t* r_full = t_empty + 2;                 // [2] residual tile of epilogue warpgroup e has landed in its staging
    uint64_t* pa_ready = r_full + 2;                // [2]  pair mode, leader: the peer's a_ready / b_full / t_empty, relayed
    uint64_t* pb_full = pa_ready + A_STAGES;     // [NB]
    uint64_t* pt_empty = pb_full + L::NB;           // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(pt_empty + 2);
    static_assert((3 * A_STAGES + 3 * L::NB + 8) * 8 + 4 <= 512, "barrier region");
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int a_bytes = K1 ? 256 * 128 : (p.imgs_per_super == 1 ? 34 * 10 * 128 : 2 * 18 * 10 * 128);

    pdl_launch_dependents();
    if (warp == 0 && lane == 0) {
        prefetch_tmap(&maps.a[0]); prefetch_tmap(&maps.b); prefetch_tmap(&maps.out[0]);
        for (int s = 0; s < A_STAGES; ++s) { mbar_init(&a_full[s], 1); mbar_init(&a_empty[s], 1); mbar_init(&a_ready[s], V2_XF_THREADS); }
        for (int s = 0; s < L::NB; ++s) { mbar_init(&b_full[s], 1); mbar_init(&b_empty[s], 1); }
        for (int s = 0; s < 2; ++s) { mbar_init(&t_full[s], 1); mbar_init(&t_empty[s], 256); mbar_init(&r_full[s], 1); }
        for (int s = 0; s < A_STAGES; ++s) mbar_init(&pa_ready[s], 1);
        for (int s = 0; s < L::NB; ++s) mbar_init(&pb_full[s], 1);
        for (int s = 0; s < 2; ++s) mbar_init(&pt_empty[s], 1);
        fence_barrier_init();
    }
    if (warp == 1) { if (PAIR) tmem_alloc_pair<512>(tmem_slot); else tmem_alloc<512>(tmem_slot); }
    tc_fence_before();
    if (PAIR) cluster_sync_all(); else __syncthreads();      // pair: both CTAs' barriers exist before any remote arrive
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    pdl_wait();                                              // prologue done; every input comes from earlier kernels of the chain

    if (warp == 0) {
        if (elect_one()) {
            // ===================== TMA producer, activations =====================
            // (the weight tiles have their own producer thread, warp 2: one thread issuing both streams in program order would
            // reach the halo load of chunk i+1 only after the last four weight taps of chunk i have found a free slot, i.e. at
            // 5/9 of chunk i's MMAs -- measured: the MMA issuer then waits 38% of its time for a_ready on the N = 128 layers)
            int as = 0; uint32_t aph = 0;
            V2_TL_DECL;
            for (int it = 0, w; (w = v2_item(p, it)) >= 0 && !V2_EXP(64); ++it) {
                const V2Work wk = v2_decode(p, w);
                for (int sg = 0; sg < p.n_seg; ++sg) {
                    const V2Seg sp = p.seg[sg];
                    for (int ch = 0; ch < sp.chunks; ++ch) {
                        V2_TW(0, mbar_wait(&a_empty[as], aph ^ 1u));
                        if (V2_EXP(16) && (it > 0 || aph)) { mbar_arrive(&a_full[as]); }
                        else {
                            mbar_arrive_expect_tx(&a_full[as], a_bytes);
                            tma_load_4d(smem + as * L::A_SLOT, &maps.a[sg], &a_full[as], ch * 64, wk.x0 - (K1 ? 0 : 1), wk.y0 - (K1 ? 0 : 1), wk.n0);
                        }
                        if (++as == A_STAGES) { as = 0; aph ^= 1u; }
                    }
                }
            }
            V2_TL_FLUSH(0, 2);
        }
    } else if (warp == 2) {
        if (elect_one()) {
            // ===================== TMA producer, weights =====================
            int bs = 0; uint32_t bph = 0;
            V2_TL_DECL;
            if (RES && !V2_EXP(64)) {                           // whole weight matrix, once (n_ntiles == 1)
                int nkb = 0;
                for (int sg = 0; sg < p.n_seg; ++sg) nkb += p.seg[sg].chunks * p.seg[sg].taps;
                mbar_arrive_expect_tx(&b_full[0], nkb * L::B_TILE);
                for (int kb = 0; kb < nkb; ++kb)
                    tma_load_2d(smem + L::OFF_B + kb * L::B_TILE, &maps.b, &b_full[0], kb * 64, 0);
            }
            for (int it = 0, w; !RES && (w = v2_item(p, it)) >= 0 && !V2_EXP(64); ++it) {
                const V2Work wk = v2_decode(p, w);
                for (int sg = 0; sg < p.n_seg; ++sg) {
                    const V2Seg sp = p.seg[sg];
                    for (int ch = 0; ch < sp.chunks; ++ch) {
                        for (int tap = 0; tap < sp.taps; ++tap) {
                            const int kb = sp.kb_base + tap * sp.kb_stride + ch;
                            V2_TW(1, mbar_wait(&b_empty[bs], bph ^ 1u));
                            if (V2_EXP(1) && (it > 0 || bph)) { mbar_arrive(&b_full[bs]); }
                            else if (PAIR) {                             // this CTA's half of the weight tile (maps.b box: BN/2 rows)
                                mbar_arrive_expect_tx(&b_full[bs], L::B_TILE / 2);
                                tma_load_2d(smem + L::OFF_B + bs * L::B_TILE, &maps.b, &b_full[bs], kb * 64,
                                            wk.nt * BN + (int)rank * (BN / 2));
                            } else {
                                mbar_arrive_expect_tx(&b_full[bs], L::B_TILE);
                                tma_load_2d(smem + L::OFF_B + bs * L::B_TILE, &maps.b, &b_full[bs], kb * 64, wk.nt * BN);
                            }
                            if (++bs == L::NB) { bs = 0; bph ^= 1u; }
                        }
                    }
                }
            }
            V2_TL_FLUSH(13, 2);
        }
    } else if (warp == 1) {
        if (PAIR && rank == 1) {
            if (elect_one()) {
                // ===================== relay (peer CTA of a pair): forward local readiness to the leader =====================
                int as = 0; uint32_t aph = 0; int bs = 0; uint32_t bph = 0; int tb = 0; uint32_t tph = 0;
                for (int it = 0, w; (w = v2_item(p, it)) >= 0; ++it) {
                    mbar_wait(&t_empty[tb], tph ^ 1u);
                    mbar_arrive_remote(&pt_empty[tb], 0);
                    for (int sg = 0; sg < p.n_seg && !V2_EXP(64); ++sg) {
                        const V2Seg sp = p.seg[sg];
                        for (int ch = 0; ch < sp.chunks; ++ch) {
                            mbar_wait(&a_ready[as], aph);
                            mbar_arrive_remote(&pa_ready[as], 0);
                            for (int tap = 0; tap < sp.taps; ++tap) {
                                mbar_wait(&b_full[bs], bph);
                                mbar_arrive_remote(&pb_full[bs], 0);
                                if (++bs == L::NB) { bs = 0; bph ^= 1u; }
                            }
                            if (++as == A_STAGES) { as = 0; aph ^= 1u; }
                        }
                    }
                    if (++tb == L::NBUF) { tb = 0; tph ^= 1u; }
                }
            }
        } else if (elect_one()) {
            // ===================== MMA issuer =====================
            constexpr uint32_t idesc = make_idesc_bf16(PAIR ? 256 : 128, BN);
            auto mma = [&](uint32_t d, uint64_t da, uint64_t db, uint32_t acc) {
                if (PAIR) umma_bf16_pair(d, da, db, idesc, acc); else umma_bf16(d, da, db, idesc, acc);
            };
            auto commit = [&](uint64_t* bar) { if (PAIR) umma_commit_pair(bar); else umma_commit(bar); };
            int as = 0; uint32_t aph = 0; int bs = 0; uint32_t bph = 0; int tb = 0; uint32_t tph = 0;
            V2_TL_DECL;
            if (RES && !V2_EXP(64)) mbar_wait(&b_full[0], 0);
            for (int it = 0, w; (w = v2_item(p, it)) >= 0; ++it) {
                const V2Work wk = v2_decode(p, w);
                V2_TW(0, mbar_wait(&t_empty[tb], tph ^ 1u));          // epilogue drained this accumulator pair
                if (PAIR) mbar_wait_cluster(&pt_empty[tb], tph);      // ... and so did the peer's
                tc_fence_after();
                uint32_t first = 1;
                for (int sg = 0; sg < p.n_seg; ++sg) {
                    const V2Seg sp = p.seg[sg];
                    for (int ch = 0; ch < sp.chunks; ++ch) {
                        if (!V2_EXP(64)) V2_TW(1, mbar_wait(&a_ready[as], aph));   // landed AND transformed
                        if (PAIR && !V2_EXP(64)) mbar_wait_cluster(&pa_ready[as], aph);
                        tc_fence_after();
                        const uint32_t a_base = smem_u32(smem + as * L::A_SLOT);
                        for (int tap = 0; tap < sp.taps; ++tap) {
                            // 9 taps: 3x3 window; 4 taps: the 2x2 window of sub-pixel phase (py, px) of a fused nearest-2x
                            // upsample (rows y+py-1, y+py of the low-res input); 1 tap: centre
                            const int dy = sp.taps == 9 ? tap / 3 : (sp.taps == 4 ? (wk.phase >> 1) + (tap >> 1) : 1);
                            const int dx = sp.taps == 9 ? tap % 3 : (sp.taps == 4 ? (wk.phase & 1) + (tap & 1) : 1);
                            if (RES) bs = sp.kb_base + tap * sp.kb_stride + ch;
                            else if (!V2_EXP(64)) {
                                V2_TW(2, mbar_wait(&b_full[bs], bph));
                                if (PAIR) mbar_wait_cluster(&pb_full[bs], bph);   // (skipped with the b_full wait under mask 64)
                                tc_fence_after();
                            }
                            const uint64_t db = make_smem_desc_sw128(smem_u32(smem + L::OFF_B + bs * L::B_TILE));
                            // the K steps of the two M tiles alternate, so that consecutive MMAs accumulate into different TMEM tiles
                            uint64_t da[V2_MT];
#pragma unroll
                            for (int mt = 0; mt < V2_MT; ++mt)
                                da[mt] = K1 ? make_smem_desc_sw128(a_base + mt * 16384, 1024)      // rows mt*128 .. +127 of the box
                                            : make_smem_desc_sw128(a_base + ((mt * p.row_off + dy) * 10 + dx) * 128, 1280);
#pragma unroll
                            for (int k = 0; k < 4; ++k) {
#pragma unroll
                                for (int mt = 0; mt < V2_MT; ++mt)
                                    mma(tmem + (tb * V2_MT + mt) * BN, da[mt] + 2 * k, db + 2 * k, (first && k == 0) ? 0u : 1u);
                            }
                            first = 0;
                            if (!RES) {
                                commit(&b_empty[bs]);
                                if (++bs == L::NB) { bs = 0; bph ^= 1u; }
                            }
                        }
                        commit(&a_empty[as]);
                        if (++as == A_STAGES) { as = 0; aph ^= 1u; }
                    }
                }
                commit(&t_full[tb]);
                if (++tb == L::NBUF) { tb = 0; tph ^= 1u; }
            }
            V2_TL_FLUSH(3, 3);
        }
    } else if (warp >= V2_XF_BASE / 32) {
        // ===================== input transform (warps 12..19): GroupNorm affine (+SiLU) in place =====================
        const int tt = threadIdx.x - V2_XF_BASE, lv = tt & 7, r0 = tt >> 3;      // 8-channel vector, first row
        const int rows_per_img = p.imgs_per_super == 1 ? 340 : 180;
        const int hy0 = r0 / 10, hx0 = r0 - hy0 * 10;                    // halo coordinates of the first row (10 pixels per halo row)
        constexpr int RSTEP = V2_XF_THREADS / 8;                         // rows advance by 32 = 3 halo rows + 2 pixels
        int as = 0; uint32_t aph = 0;
        V2_TL_DECL;
        for (int it = 0, w; (w = v2_item(p, it)) >= 0 && !V2_EXP(64); ++it) {
            const V2Work wk = v2_decode(p, w);
            for (int sg = 0; sg < p.n_seg; ++sg) {
                const V2Seg sp = p.seg[sg];
                for (int ch = 0; ch < sp.chunks; ++ch) {
                    float sc[8], sh[8];                                    // image 0 of the super-tile: in flight while the TMA lands
                    const float pre = sp.xform == 2 ? 0.5f : 1.0f;         // SiLU works on h = y/2: silu(y) = h + h*tanh(h)
                    auto load_ss = [&](int im) {
                        const float4* ss = reinterpret_cast<const float4*>(
                            p.gn_ss + (size_t)(wk.n0 + im) * p.gn_C + sp.ss_off + ch * 64 + lv * 8);
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const float4 t = __ldg(ss + j);
                            sc[2 * j] = pre * t.x; sh[2 * j] = pre * t.y; sc[2 * j + 1] = pre * t.z; sh[2 * j + 1] = pre * t.w;
                        }
                    };
                    if (sp.xform) load_ss(0);
                    V2_TW(0, mbar_wait(&a_full[as], aph));
#ifdef SYNT_EXPERIMENTS
                    const long long tx0 = clock64();
#endif
                    if (sp.xform && !V2_EXP(2)) {
                        uint8_t* slot = smem + as * L::A_SLOT;
                        auto run = [&](auto silu_tag) {
                            constexpr bool SILU = decltype(silu_tag)::value;
                            auto apply = [&](uint4& v) {                  // 8 channels of one pixel: affine (+SiLU) in fp32
                                __nv_bfloat162* h2 = reinterpret_cast<__nv_bfloat162*>(&v);
#pragma unroll
                                for (int j = 0; j < 4; ++j) {
                                    if (V2_EXP(128)) continue;             // experiment: copy only
                                    float2 f = __bfloat1622float2(h2[j]);
                                    f.x = fmaf(f.x, sc[2 * j], sh[2 * j]);
                                    f.y = fmaf(f.y, sc[2 * j + 1], sh[2 * j + 1]);
                                    if (SILU && !V2_EXP(512)) {
                                        float t0, t1;
                                        asm("tanh.approx.f32 %0, %1;" : "=f"(t0) : "f"(f.x));
                                        asm("tanh.approx.f32 %0, %1;" : "=f"(t1) : "f"(f.y));
                                        f.x = fmaf(f.x, t0, f.x); f.y = fmaf(f.y, t1, f.y);
                                    }
                                    h2[j] = __floats2bfloat162_rn(f.x, f.y);
                                }
                            };
#pragma unroll
                            for (int im = 0; im < 2; ++im) {
                                if (im >= p.imgs_per_super || wk.n0 + im >= p.B) continue;
                                if (im == 1) load_ss(1);                   // 16x16 layers: second image of the super-tile
                                constexpr int UR = 4;                      // rows in flight per thread (hides LDS/MUFU latency)
                                if (K1) {                                  // no halo: rows r0, r0 + 32, ... of this image's pixels
                                    const int rpi = p.imgs_per_super == 1 ? 256 : 128;
                                    for (int k0 = 0; k0 < rpi / RSTEP; k0 += UR) {
                                        uint4 v[UR]; uint4* ptr[UR];
#pragma unroll
                                        for (int uu = 0; uu < UR; ++uu) {
                                            const int r = im * rpi + r0 + (k0 + uu) * RSTEP;
                                            ptr[uu] = reinterpret_cast<uint4*>(slot + r * 128 + ((lv ^ (r & 7)) << 4));
                                            v[uu] = *ptr[uu];
                                        }
#pragma unroll
                                        for (int uu = 0; uu < UR; ++uu) { apply(v[uu]); *ptr[uu] = v[uu]; }
                                    }
                                    continue;
                                }
                                int hy = wk.y0 - 1 + hy0, hx = wk.x0 - 1 + hx0;     // image coordinates of the current row
                                for (int rr = r0; rr < rows_per_img;) {
                                    uint4 v[UR]; uint4* ptr[UR]; bool ok[UR];
#pragma unroll
                                    for (int uu = 0; uu < UR; ++uu) {
                                        // out-of-image halo pixels stay exactly 0 (= the padding of the activated tensor)
                                        ok[uu] = rr < rows_per_img && (unsigned)hy < (unsigned)p.H && (unsigned)hx < (unsigned)p.W;
                                        const int r = im * 180 + rr;
                                        ptr[uu] = reinterpret_cast<uint4*>(slot + r * 128 + ((lv ^ (r & 7)) << 4));
                                        if (ok[uu]) v[uu] = *ptr[uu];
                                        rr += RSTEP; hy += 3; hx += 2;
                                        if (hx >= wk.x0 + 9) { hx -= 10; ++hy; }
                                    }
#pragma unroll
                                    for (int uu = 0; uu < UR; ++uu) {
                                        if (!ok[uu]) continue;
                                        apply(v[uu]);
                                        if (!V2_EXP(256) || v[uu].x == 0x12345678u) *ptr[uu] = v[uu];
                                    }
                                }
                            }
                        };
                        if (sp.xform == 2) run(std::true_type{}); else run(std::false_type{});
                        fence_proxy_async();                               // generic-proxy writes -> UMMA (async proxy)
                    }
                    mbar_arrive(&a_ready[as]);
#ifdef SYNT_EXPERIMENTS
                    tlacc[1] += clock64() - tx0; tlacc[2] += 1;
#endif
                    if (++as == A_STAGES) { as = 0; aph ^= 1u; }
                }
            }
        }
        if (tt == 0) V2_TL_FLUSH(7, 3);
    } else if (warp >= V2_EPI_BASE / 32) {
        // ===================== epilogue: warpgroup e (warps 4..7 / 8..11) owns M tile e of every super-tile ==========
        const int e = (warp - V2_EPI_BASE / 32) >> 2;                // warpgroup = M tile index
        const int q = warp & 3, r = q * 32 + lane;                  // accumulator row = pixel (r/8, r%8) of the 16x8 tile
        constexpr int EC = L::EC, NPASS = L::NPASS;
        const int et = (threadIdx.x - V2_EPI_BASE) & 127;           // thread index within the warpgroup
        uint8_t* staging = smem + L::OFF_STAGING + e * L::STAGING;
        float* bias_s = reinterpret_cast<float*>(smem + L::OFF_BIAS) + e * BN;
        const int sw = r & 7;
        const bool lead_warp = q == 0;                              // its elected lane owns this warpgroup's TMA loads/stores
        const uint32_t bar_id = 1 + e;
        auto wg_sync = [&]() { asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory"); };
        // tile e of work item w: image, first row, validity (B odd with two images per super-tile)
        auto tile_of = [&](const V2Work& wk, int& n_img, int& ty0) {
            n_img = wk.n0 + (p.imgs_per_super == 1 ? 0 : e);
            ty0 = wk.y0 + (p.imgs_per_super == 1 ? 16 * e : 0);
            return n_img < p.B;                                      // H, W are multiples of the tile
        };
        auto load_residual = [&](const V2Work& wk, int pass) {     // residual tile (one pass) -> staging, same swizzled layout
            int n_img, ty0;
            if (!tile_of(wk, n_img, ty0)) return;
            mbar_arrive_expect_tx(&r_full[e], L::STAGING);
#pragma unroll
            for (int j = 0; j < EC / 64; ++j)
                tma_load_4d(staging + j * 16384, &maps.res, &r_full[e], wk.ntr * BN + pass * EC + j * 64, wk.x0, ty0, n_img);
        };
        int tb = 0; uint32_t tph = 0, rph = 0; int last_nt = -1;
        float acc[8];                                               // GroupNorm partials (4 columns x (sum, sumsq)) carried across tiles
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i] = 0.f;
        if (p.has_res && lead_warp) {
            const int w0 = v2_item(p, 0);
            if (w0 >= 0) { if (elect_one()) load_residual(v2_decode(p, w0), 0); }
        }
        V2_TL_DECL;
        for (int it = 0, w; (w = v2_item(p, it)) >= 0; ++it) {
            const V2Work wk = v2_decode(p, w);
            int n_img, ty0;
            const bool valid = tile_of(wk, n_img, ty0);
            if (wk.ntr != last_nt) {                                  // (bias + time-embedding row) of this N tile -> smem
                wg_sync();
                for (int i = et; i < BN; i += 128) bias_s[i] = p.bias[wk.ntr * BN + i] + (p.bias2 ? p.bias2[wk.ntr * BN + i] : 0.f);
                last_nt = wk.ntr;
            }
            V2_TW(0, mbar_wait(&t_full[tb], tph));
            tc_fence_after();
            const bool has_res = p.has_res && valid;
#pragma unroll 1
            for (int pass = 0; pass < NPASS; ++pass) {
                const int cb = pass * EC;                            // first column of this pass within the N tile
                if (has_res) { mbar_wait(&r_full[e], rph); rph ^= 1u; }  // landed; also means the staging tile was free
                wg_sync();                                           // staging free (leader waited for the previous store), bias visible
#pragma unroll
                for (int c0 = 0; c0 < EC; c0 += 32) {
                    if (V2_EXP(32)) break;
                    uint32_t v[32];
                    tmem_ld_32x32b_x32(tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)((tb * V2_MT + e) * BN + cb + c0), v);
                    tmem_ld_wait();
                    uint8_t* srow = staging + (c0 >> 6) * 16384 + r * 128;
#pragma unroll
                    for (int g = 0; g < 4; ++g) {
                        float f[8];
                        const float4 b0 = *reinterpret_cast<const float4*>(bias_s + cb + c0 + g * 8);
                        const float4 b1 = *reinterpret_cast<const float4*>(bias_s + cb + c0 + g * 8 + 4);
                        f[0] = __uint_as_float(v[g * 8 + 0]) + b0.x; f[1] = __uint_as_float(v[g * 8 + 1]) + b0.y;
                        f[2] = __uint_as_float(v[g * 8 + 2]) + b0.z; f[3] = __uint_as_float(v[g * 8 + 3]) + b0.w;
                        f[4] = __uint_as_float(v[g * 8 + 4]) + b1.x; f[5] = __uint_as_float(v[g * 8 + 5]) + b1.y;
                        f[6] = __uint_as_float(v[g * 8 + 6]) + b1.z; f[7] = __uint_as_float(v[g * 8 + 7]) + b1.w;
                        uint4* slot = reinterpret_cast<uint4*>(srow + (((((c0 & 63) >> 3) + g) ^ sw) << 4));
                        if (has_res) {                               // in-place: this thread's own 16 bytes of the residual tile
                            const uint4 rv = *slot;
                            const __nv_bfloat162* rh = reinterpret_cast<const __nv_bfloat162*>(&rv);
#pragma unroll
                            for (int j = 0; j < 4; ++j) {
                                const float2 t = __bfloat1622float2(rh[j]);
                                f[2 * j] += t.x; f[2 * j + 1] += t.y;
                            }
                        }
                        if (p.relu) {
#pragma unroll
                            for (int j = 0; j < 8; ++j) f[j] = fmaxf(f[j], 0.f);
                        }
                        uint4 pk;
                        __nv_bfloat162* h2 = reinterpret_cast<__nv_bfloat162*>(&pk);
#pragma unroll
                        for (int j = 0; j < 4; ++j) h2[j] = __floats2bfloat162_rn(f[2 * j], f[2 * j + 1]);
                        *slot = pk;
                    }
                }
                if (pass == NPASS - 1) {
                    tc_fence_before();
                    mbar_arrive(&t_empty[tb]);                       // this warpgroup's accumulator tile is fully read
                }
                fence_proxy_async();
                wg_sync();
                if (lead_warp && valid && !V2_EXP(8)) {
                    if (elect_one()) {
#pragma unroll
                        for (int j = 0; j < EC / 64; ++j)
                            tma_store_4d(&maps.out[wk.phase], staging + j * 16384, wk.ntr * BN + cb + j * 64, wk.x0, ty0, n_img);
                        tma_store_commit();
                    }
                }
                if (p.stats && !V2_EXP(4)) {
                    // fused GroupNorm statistics: per-channel (sum, sumsq) of the bf16 values just staged.  Warp q of the
                    // warpgroup owns EC/4 columns; lane = (column quad cq, row part rp): 8-byte loads of 4 columns, the row
                    // parts interleaved so that the lanes of one load phase hit distinct banks of the swizzled tile; the
                    // partials stay in registers across the tiles of a chunk (single-pass tiles), fixed summation order.
                    constexpr int CQW = EC / 16, RP = 32 / CQW, SUB = 16 / RP;     // EC=128: 8 quads x 4 parts x 4-row runs
                    const int cq = lane % CQW, rp = lane / CQW;
                    const int col0 = q * (EC / 4) + cq * 4;
                    if (valid) {
                        const uint8_t* sb = staging + (col0 >> 6) * 16384 + (col0 & 7) * 2;
                        const int ck = (col0 & 63) >> 3;
#pragma unroll
                        for (int k = 0; k < 8; ++k) {
#pragma unroll
                            for (int j = 0; j < SUB; ++j) {
                                const int row = (2 * k + rp / (RP / 2)) * 8 + (rp % (RP / 2)) * SUB + j;
                                const uint2 u = *reinterpret_cast<const uint2*>(sb + row * 128 + ((ck ^ (row & 7)) << 4));
                                const float x0 = __uint_as_float(u.x << 16), x1 = __uint_as_float(u.x & 0xffff0000u);
                                const float x2 = __uint_as_float(u.y << 16), x3 = __uint_as_float(u.y & 0xffff0000u);
                                acc[0] += x0; acc[1] = fmaf(x0, x0, acc[1]); acc[2] += x1; acc[3] = fmaf(x1, x1, acc[3]);
                                acc[4] += x2; acc[5] = fmaf(x2, x2, acc[5]); acc[6] += x3; acc[7] = fmaf(x3, x3, acc[7]);
                            }
                        }
                    }
                    // one partial row per tile (16x16 layers, multi-pass tiles) or per chunk (chunks never straddle an
                    // (image, N tile)) and epilogue warpgroup
                    const bool per_tile = p.imgs_per_super == 2 || NPASS > 1 || PAIR;
                    const bool flush = per_tile ? valid : (w + 1) % p.chunk == 0;
                    if (flush) {
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
#pragma unroll
                            for (int o = CQW; o < 32; o <<= 1) acc[i] += __shfl_xor_sync(0xffffffffu, acc[i], o);
                        }
                        if (rp == 0) {
                            const int per_img = p.tiles_x * p.supers_per_img;
                            const int slot = p.imgs_per_super == 2 ? wk.phase * p.tiles_x + (wk.x0 >> 3)
                                           : (NPASS > 1 || PAIR) ? (wk.phase * per_img + w % per_img) * V2_MT + e
                                                       : (wk.phase * (per_img / p.chunk) + (w % per_img) / p.chunk) * V2_MT + e;
                            float4* dst = reinterpret_cast<float4*>(
                                p.stats + ((size_t)(p.imgs_per_super == 2 ? n_img : wk.n0) * p.stats_slots + slot) * p.Cout +
                                wk.ntr * BN + cb + col0);
                            dst[0] = make_float4(acc[0], acc[1], acc[2], acc[3]);
                            dst[1] = make_float4(acc[4], acc[5], acc[6], acc[7]);
                        }
#pragma unroll
                        for (int i = 0; i < 8; ++i) acc[i] = 0.f;
                    }
                }
                wg_sync();                                           // every statistics read of the staging tile is done
                if (lead_warp) {
                    if (elect_one()) {
                        tma_store_wait_read();                       // ... and so is the TMA store's: the staging tile is free
                        if (p.has_res) {
                            if (pass + 1 < NPASS) load_residual(wk, pass + 1);
                            else {
                                const int wn = v2_item(p, it + 1);
                                if (wn >= 0) load_residual(v2_decode(p, wn), 0);
                            }
                        }
                    }
                }
            }
            if (++tb == L::NBUF) { tb = 0; tph ^= 1u; }
        }
        if (et == 0 && e == 0) V2_TL_FLUSH(11, 1);
        if (lead_warp) { if (elect_one()) tma_store_wait_all(); }
    }
    tc_fence_before();
    if (PAIR) cluster_sync_all(); else __syncthreads();      // pair: neither CTA may leave while the other's MMAs read its smem
    if (warp == 1) { if (PAIR) tmem_dealloc_pair<512>(tmem); else tmem_dealloc<512>(tmem); }
    (void)out;
}

bool conv_tc2_supported(const ConvArgs& a) {
    if (a.up2x && (a.KH != 3 || a.Cin1 || a.sc0_C || a.sc1_C || a.gn_mode || a.residual)) return false;
    const bool k3 = a.KH == 3 && a.KW == 3 && a.pad == 1;
    const bool k1 = a.KH == 1 && a.KW == 1 && a.pad == 0 && a.sc0_C == 0 && a.sc1_C == 0 && a.Cin1 == 0;   // centre-tap-only conv
    if (!(k3 || k1) || a.stride != 1 || a.sc_stride != 1) return false;
    if (a.Cin % 64 || a.Cin1 % 64 || a.sc0_C % 64 || a.sc1_C % 64 || a.Cout % 64) return false;
    if (a.W % 8 == 0 && (a.H % 32 == 0 || a.H == 16)) return true;
    // ragged planes (ResNet18: 56x56, 28x28): the tile grid is rounded up, TMA zero-fills the loads and clips the stores
    // beyond the plane.  Not with fused statistics / transformed input / upsampling (out-of-plane pixels would count).
    // planes of 9..16 rows (ResNet18 14x14) use the two-images-per-super-tile mode of the 16x16 layers
    return a.H > 8 && a.W >= 8 && a.stats_out == nullptr && a.gn_mode == 0 && !a.up2x;
}
static inline bool v2_two_img(const ConvArgs& a) { return a.H <= 16; }      // one 16-row M tile per image, two images per super-tile
static inline int v2_tiles_x(const ConvArgs& a) { return (a.W + 7) / 8; }
static inline int v2_supers(const ConvArgs& a) { return v2_two_img(a) ? 1 : (a.H + 31) / 32; }

// N tile: 256 for the Cout = 256 3x3 convolutions at 32x32 and above (K >= 1152: the un-overlapped epilogue of the
// single-buffered accumulators stays below ~10% of the mainloop; measured 10-18% faster than two N = 128 tiles),
// else 128, else 64.  The 16x16 layers keep N = 128: with 128 work items on 148 SMs the longer, un-overlapped items
// measured 30% slower.
static int v2_bn(const ConvArgs& a) {
    if (a.Cout % 256 == 0 && a.KH == 3 && !a.up2x && a.H >= 32 && 9 * (a.Cin + a.Cin1) >= 1152) {
        static const char* e = getenv("SYNT_CONV_BN256");
        if (!(e && e[0] == '0')) return 256;
    }
    return a.Cout % 128 == 0 ? 128 : 64;
}

// CTA-pair mode (cta_group::2): N = 128 tiles with streamed weights and an even number of super-tiles per image
// Off by default: functionally complete (tests run it), but measured SLOWER than the single-CTA kernel (conv total 5.4 ->
// 6.3 ms/step): the pure MMA rate only improves from 87 to 81 clk per N = 128 K-step (the per-instruction cost is
// N/2 + ~21 clk, not operand bandwidth) while the relayed barriers lengthen every hand-shake.  SYNT_CONV_PAIR=1 or
// synt_debug_set_conv_pair(1) switches it on.
static int g_conv_pair = -1;
void conv_tc2_set_pair(int on) { g_conv_pair = on; }
static bool v2_pair(const ConvArgs& a) {
    if (g_conv_pair < 0) { const char* e = getenv("SYNT_CONV_PAIR"); g_conv_pair = (e && e[0] == '1') ? 1 : 0; }
    if (!g_conv_pair) return false;
    if (v2_bn(a) != 128) return false;
    const int per_img = v2_tiles_x(a) * v2_supers(a);
    return per_img % 2 == 0;
}

// 1x1 projections on full tile grids take the no-halo, four-slot variant (V2Smem<128, false, true>)
static bool v2_k1(const ConvArgs& a) {
    if (!(a.KH == 1 && v2_bn(a) == 128 && !a.up2x)) return false;
    if (a.residual || a.stats_out) return false;      // the out-projection (residual + statistics) measured 2 us slower with 64-column passes
    if (!(a.W % 8 == 0 && (a.H % 32 == 0 || a.H == 16))) return false;     // ragged planes keep the halo geometry (zero fill)
    static const char* e = getenv("SYNT_CONV_K1");
    return !(e && e[0] == '0');
}

bool conv_tc2_is_k1(const ConvArgs& a) { return conv_tc2_supported(a) && v2_k1(a) && !v2_pair(a); }

static int v2_num_sms() {
    static int n = 0;
    if (!n) { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev); }
    return n;
}
// consecutive work items per CTA turn: the divisor of the super-tiles per image (<= 8) with the smallest
// makespan ceil(chunks / grid) * R; ties go to the larger R (fewer GroupNorm partial rows)
static int v2_chunk(const ConvArgs& a, int BN) {
    if (v2_two_img(a)) return 1;
    const int per_img = v2_tiles_x(a) * v2_supers(a);
    const long long n_work = (long long)a.B * per_img * (a.Cout / BN) * (a.up2x ? 4 : 1);
    const long long grid = n_work < v2_num_sms() ? n_work : v2_num_sms();
    int best = 1; long long best_span = -1;
    for (int R = 8; R >= 1; R >>= 1) {
        if (per_img % R) continue;
        const long long chunks = n_work / R;
        const long long span = ((chunks + grid - 1) / grid) * R;
        if (best_span < 0 || span < best_span) { best = R; best_span = span; }
    }
    return best;
}
// partial rows per image in stats_out (every row is written exactly once by the kernel)
int conv_tc2_stats_slots(const ConvArgs& a) {
    const int BN = v2_bn(a);
    const int phases = a.up2x ? 4 : 1;
    if (v2_two_img(a)) return phases * v2_tiles_x(a);                      // one row per tile
    const int per_img = v2_tiles_x(a) * v2_supers(a);
    if (BN == 256 || v2_pair(a) || v2_k1(a)) return phases * per_img * V2_MT;   // multi-pass tiles / pair mode: one row per tile
    return phases * (per_img / v2_chunk(a, BN)) * V2_MT;                   // one row per chunk and epilogue warpgroup
}

// cluster (2,1,1) launch of the CTA-pair variant
static void launch_v2_pair(const V2Maps& maps, const V2Params& p, int grid, bf16* out, cudaStream_t s) {
    using L = V2Smem<128, false>;
    auto kern = conv_tc2_kernel<128, false, true>;
    ensure_dynamic_smem((const void*)(kern), L::TOTAL);
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(V2_THREADS); cfg.dynamicSmemBytes = L::TOTAL; cfg.stream = s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    SYNT_CUDA(cudaLaunchKernelEx(&cfg, kern, maps, p, out));
}

template <int BN, bool RES, bool K1 = false>
static void launch_v2(const V2Maps& maps, const V2Params& p, int grid, bf16* out, cudaStream_t s) {
    using L = V2Smem<BN, RES, K1>;
    ensure_dynamic_smem((const void*)(conv_tc2_kernel<BN, RES, false, K1>), L::TOTAL);
    launch_pdl<true>(conv_tc2_kernel<BN, RES, false, K1>, dim3(grid), dim3(V2_THREADS), L::TOTAL, s, maps, p, out);
}

static void make_halo_map(CUtensorMap* m, const void* base, int B, int H, int W, int C, int box_h, int box_n, int box_w = 10) {
    cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
    cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
    cuuint32_t box[4] = {64, (cuuint32_t)box_w, (cuuint32_t)box_h, (cuuint32_t)box_n};
    encode_bf16_sw128(m, base, 4, dims, strides, box, "halo activation");
}

void conv_tc2(const ConvArgs& a, cudaStream_t s) {
    SYNT_CHECK(conv_tc2_supported(a), "conv_tc2: unsupported shape");
    SYNT_CHECK(a.bias != nullptr, "conv_tc2: bias required");
    const int BN = v2_bn(a);
    const bool pair = v2_pair(a);
    V2Params p{};
    p.imgs_per_super = v2_two_img(a) ? 2 : 1;
    p.row_off = v2_two_img(a) ? 18 : 16;
    p.tiles_x = v2_tiles_x(a);
    p.supers_per_img = v2_supers(a);
    const int phases = a.up2x ? 4 : 1;
    p.nt_real = a.Cout / BN;
    p.n_ntiles = p.nt_real * phases;
    const int n_super = ceil_div(a.B, p.imgs_per_super) * p.tiles_x * p.supers_per_img;
    p.n_work = n_super * p.n_ntiles;
    const bool k1 = a.KH == 1;                          // 1x1 conv == centre-tap-only segment of the same machinery
    const int Ct = a.Cin + a.Cin1;
    const void* srcs[4] = {nullptr, nullptr, nullptr, nullptr}; int src_C[4] = {0, 0, 0, 0};
    p.n_seg = 0;
    auto add_seg = [&](const void* src, int C, int taps, int kb_base, int kb_stride, int xform, int ss_off) {
        if (!C) return;
        srcs[p.n_seg] = src; src_C[p.n_seg] = C;
        p.seg[p.n_seg++] = V2Seg{C / 64, taps, kb_base, kb_stride, xform, ss_off};
    };
    if (k1) {
        add_seg(a.in, a.Cin, 1, 0, 0, a.gn_mode, 0);
    } else if (a.up2x) {
        add_seg(a.in, a.Cin, 4, 0, a.Cin / 64, 0, 0);     // weights: [4 phases x Cout][4 taps x Cin], see pack_upsample_phases
    } else {
        add_seg(a.in, a.Cin, 9, 0, Ct / 64, a.gn_mode, 0);
        add_seg(a.in1, a.Cin1, 9, a.Cin / 64, Ct / 64, a.gn_mode, a.Cin);
        add_seg(a.sc0, a.sc0_C, 1, 9 * Ct / 64, 0, 0, 0);
        add_seg(a.sc1, a.sc1_C, 1, 9 * Ct / 64 + a.sc0_C / 64, 0, 0, 0);
    }
    SYNT_CHECK(!a.gn_mode || a.gn_ss != nullptr, "conv_tc2: gn_mode without scale/shift");
    p.gn_ss = a.gn_ss; p.gn_C = Ct;
    p.B = a.B; p.H = a.H; p.W = a.W; p.Cout = a.Cout;
    p.bias = a.bias; p.bias2 = a.bias2; p.has_res = a.residual != nullptr; p.relu = a.relu;
    p.stats = a.stats_out; p.stats_slots = conv_tc2_stats_slots(a);
    p.chunk = v2_chunk(a, BN);
    {
        auto magic = [](int d) { return (uint32_t)((1ull << 32) / (unsigned long long)d + 1ull); };
        p.mg_per_img = magic(p.tiles_x * p.supers_per_img); p.mg_ntiles = magic(p.n_ntiles); p.mg_nt_real = magic(p.nt_real);
        p.mg_tiles_x = magic(p.tiles_x); p.mg_chunk = magic(p.chunk);
        SYNT_CHECK((long long)p.n_work * (p.tiles_x * p.supers_per_img) < (1ll << 32), "conv_tc2: work-item index range");
    }
    p.exp_nob = 0; p.tl = nullptr;
#ifdef SYNT_EXPERIMENTS
    { static const char* e = getenv("SYNT_EXP_NOB"); p.exp_nob = e ? atoi(e) : 0; }
    static const char* tl_env = getenv("SYNT_CONV_TL");     // file prefix: one record file per launch (debug tools only)
    static int tl_launch = 0;
    long long* tl_dev = nullptr;
    if (tl_env) { SYNT_CUDA(cudaMalloc(&tl_dev, 148 * 16 * 8)); SYNT_CUDA(cudaMemsetAsync(tl_dev, 0, 148 * 16 * 8, s)); p.tl = tl_dev; }
#endif
    V2Maps maps;
    const bool k1v = !pair && v2_k1(a);
    const int bh = k1v ? (v2_two_img(a) ? 16 : 32) : (v2_two_img(a) ? 18 : 34), bn = v2_two_img(a) ? 2 : 1;
    for (int i = 0; i < 4; ++i) {
        if (i < p.n_seg) make_halo_map(&maps.a[i], srcs[i], a.B, a.H, a.W, src_C[i], bh, bn, k1v ? 8 : 10);
        else maps.a[i] = maps.a[0];
    }
    {
        const int kt = a.up2x ? 4 * a.Cin : a.ktot();
        cuuint64_t dims[2] = {(cuuint64_t)kt, (cuuint64_t)a.Cout * phases};
        cuuint64_t strides[1] = {(cuuint64_t)kt * 2};
        cuuint32_t box[2] = {64, (cuuint32_t)(pair ? BN / 2 : BN)};       // pair mode: each CTA loads half of the N tile
        encode_bf16_sw128(&maps.b, a.weight, 2, dims, strides, box, "v2 weight");
    }
    for (int ph = 0; ph < 4; ++ph) {
        if (ph >= phases) { maps.out[ph] = maps.out[0]; continue; }
        // phase (py, px) of a fused 2x upsample writes output pixels (2y+py, 2x+px): a strided view of `out`
        const int up = a.up2x ? 2 : 1, py = ph >> 1, px = ph & 1;
        const int Wo = a.W * up, Ho = a.H * up;
        const bf16* base = (const bf16*)a.out + ((size_t)py * Wo + px) * a.Cout;
        cuuint64_t dims[4] = {(cuuint64_t)a.Cout, (cuuint64_t)a.W, (cuuint64_t)a.H, (cuuint64_t)a.B};
        cuuint64_t strides[3] = {(cuuint64_t)up * a.Cout * 2, (cuuint64_t)up * Wo * a.Cout * 2, (cuuint64_t)Ho * Wo * a.Cout * 2};
        cuuint32_t box[4] = {64, 8, 16, 1};
        encode_bf16_sw128(&maps.out[ph], base, 4, dims, strides, box, "v2 output");
    }
    if (a.residual) {
        cuuint64_t dims[4] = {(cuuint64_t)a.Cout, (cuuint64_t)a.W, (cuuint64_t)a.H, (cuuint64_t)a.B};
        cuuint64_t strides[3] = {(cuuint64_t)a.Cout * 2, (cuuint64_t)a.W * a.Cout * 2, (cuuint64_t)a.H * a.W * a.Cout * 2};
        cuuint32_t box[4] = {64, 8, 16, 1};
        encode_bf16_sw128(&maps.res, a.residual, 4, dims, strides, box, "v2 residual");
    } else {
        maps.res = maps.out[0];
    }
    const int num_sms = v2_num_sms();
    const int grid = p.n_work < num_sms ? p.n_work : num_sms;
    if (pair) {
        const int clusters = (num_sms / 2 < p.n_work / 2) ? num_sms / 2 : p.n_work / 2;
        launch_v2_pair(maps, p, 2 * clusters, (bf16*)a.out, s);
        return;
    }

    const bool resident = BN == 64 && p.n_ntiles == 1 && !a.up2x && a.ktot() / 64 <= 12;
    if (BN == 256)     launch_v2<256, false>(maps, p, grid, (bf16*)a.out, s);
    else if (BN == 128 && k1v) launch_v2<128, false, true>(maps, p, grid, (bf16*)a.out, s);
    else if (BN == 128) launch_v2<128, false>(maps, p, grid, (bf16*)a.out, s);
    else if (resident) launch_v2<64, true>(maps, p, grid, (bf16*)a.out, s);
    else               launch_v2<64, false>(maps, p, grid, (bf16*)a.out, s);
#ifdef SYNT_EXPERIMENTS
    if (tl_dev) {                                          // averages over the CTAs of this launch, cycles
        std::vector<long long> h(148 * 16);
        SYNT_CUDA(cudaStreamSynchronize(s));
        SYNT_CUDA(cudaMemcpy(h.data(), tl_dev, h.size() * 8, cudaMemcpyDeviceToHost));
        cudaFree(tl_dev);
        double avg[16] = {0};
        for (int c = 0; c < grid; ++c) for (int i = 0; i < 16; ++i) avg[i] += (double)h[c * 16 + i] / grid;
        fprintf(stderr, "[conv_tl %d] M=%d N=%d(BN %d) K=%d gn=%d | producer: wait a_empty %.0f b_empty %.0f total %.0f | mma: wait t_empty %.0f "
                "a_ready %.0f b_full %.0f total %.0f | xform: wait a_full %.0f busy %.0f chunks %.0f total %.0f | epi: wait t_full %.0f total %.0f\n",
                tl_launch++, a.B * a.H * a.W, a.Cout, BN, a.ktot(), a.gn_mode, avg[0], avg[14], avg[2], avg[3], avg[4], avg[5], avg[6], avg[7],
                avg[8], avg[9], avg[10], avg[11], avg[12]);
    }
#endif
}

}  // namespace synt
