// Fused tcgen05 self-attention for the UNet's low-resolution blocks (diffusers Attention +
// AttnProcessor2_0 reached from core/generator/image_generator.py:400): 32 heads of d = 8,
// sequence N = 1024 (32x32) or 256 (16x16).
//
// Inputs:
//   qkv  [B*N, 768] bf16 = ( q 256 | k 256 | v 256 ), the plain output of the fused q/k/v projection; q is pre-scaled by
//          log2(e)/sqrt(8) (folded into the projection weights).  A head is 8 columns, UMMA_K is 16 for bf16, so one
//          K step covers a PAIR of heads: the key operand is the natural 16-column pair (k_2p | k_2p+1) and the query
//          operand of head j is (q_j | 0) for even j, (0 | q_j) for odd j -- a zero-masked copy of the 128-query tile
//          that the softmax warps build once per CTA in shared memory.  No padded tensor exists in HBM.
//   The P.V operand of a head is V^T padded to 16 rows: rows 0..7 = v dims, row 8 = 1 (so the P.V MMA also produces the
//          softmax denominator), rows 9..15 = 0.  Warp 3 builds it per 64-key stage directly in shared memory from the v
//          columns of qkv (registers -> swizzled 32-bit stores, next stage prefetched); no V^T tensor exists in HBM.
//
// One CTA = 128 queries x 4 heads of one image, TWO CTAs resident per SM; one pass over the keys in chunks of 64:
//   S  = Q_h K_h^T     one tcgen05.mma  M128 x N64 x K16  -> TMEM (fp32, log2 domain), 3 S buffers
//   P  = exp2(S - m)   softmax warps: tcgen05.ld -> ex2 -> bf16 -> swizzled st.shared
//   O_h += P V_h       four tcgen05.mma M128 x N16 x K16, P from shared memory
// Online softmax with LAZY rescaling: the running reference m of a row only moves when the chunk
// maximum exceeds it by more than 2^8; then O_h (16 TMEM columns incl. the denominator column)
// is scaled in place by the softmax warps (tcgen05.ld/st) while no MMA on O_h is in flight.  The
// result is mathematically the exact softmax.  d = 8 makes the kernel MUFU bound (N^2 ex2 per head,
// 16 ex2/clk/SM; the MMAs are <10% of its time): the design goal is to keep the MUFU pipe busy, hence
// two co-resident CTAs (256 TMEM columns, 64 KB smem, 80 registers/thread at launch, re-split 48 / 96 by setmaxnreg) = 4 softmax warps per SM
// sub-partition whose load/max/exp phases interleave, and three rotating P tiles (unit u -> tile u % 3, like the
// S buffers) so that a softmax warpgroup does not wait for the P.V MMA of its previous unit.
//
// Warp roles (384 threads): warp 0 TMA producer, warp 1 TMEM allocator + S-MMA issuer, warp 2
// P.V-MMA issuer, warp 3 V^T builder, warps 4..7 softmax warpgroup 0 (heads 0,2), warps 8..11 softmax
// warpgroup 1 (heads 1,3).  The softmax warps keep only 32 S values live (<= 80 registers/thread).
#include "kernels.cuh"
#include "ptx.cuh"
#include "tmap.cuh"
#include <cstdio>
#include <mutex>
#include <type_traits>
#include <vector>

namespace synt {

using namespace ptx;

constexpr int ATC_THREADS = 384;
#ifndef ATC_STAGES_N
#define ATC_STAGES_N 3
#endif
constexpr int ATC_STAGES = ATC_STAGES_N;
constexpr int ATC_KEYS = 64;                           // keys per chunk
constexpr int ATC_NS = 3;                              // S buffers (64 TMEM columns each) and P tiles
constexpr int ATC_Q_BYTES = 128 * 128;                 // 128 queries x (4 heads x 16) bf16, zero-masked
constexpr int ATC_K_BYTES = ATC_KEYS * 128;            // 64 keys x 8 heads x 8 dims bf16 (natural layout, this CTA uses 4 heads)
constexpr int ATC_V_BYTES = 4 * 16 * 128;              // 4 heads x 16 rows x 64 keys
constexpr int ATC_STAGE_BYTES = ATC_K_BYTES + ATC_V_BYTES;
// P through TENSOR MEMORY (round 2): the softmax warps write the bf16 probabilities with tcgen05.st over the S columns they
// have just consumed (unit u: P in columns [64 u', 64 u' + 32) of its S buffer, two keys per 32-bit column) and the P.V MMAs
// take their A operand from there (tcgen05.mma [d], [a_tmem], b_desc; layout checked by tools/ubench/mma_ts.cu).  No P tile
// in shared memory: -32 KB of shared-memory traffic per unit (16 KB of stores + 16 KB of operand fetch), no swizzled store
// addresses, no generic->async proxy fence in the unit body.  The S buffer is released by the P.V issuer's commit instead of
// by the softmax warps' last load.  ATC_P_TMEM=0 keeps the shared-memory path (three rotating P tiles).
#ifndef ATC_P_TMEM
#define ATC_P_TMEM 1
#endif
constexpr int ATC_P_BYTES = ATC_P_TMEM ? 0 : 128 * 128; // 128 queries x 64 keys bf16
constexpr int ATC_OFF_STAGE = ATC_Q_BYTES;
constexpr int ATC_OFF_P = ATC_OFF_STAGE + ATC_STAGES * ATC_STAGE_BYTES;
constexpr int ATC_OFF_BAR = ATC_OFF_P + ATC_NS * ATC_P_BYTES;   // P tiles rotate with the S buffers (unit u -> u % 3)
constexpr int ATC_SMEM = ATC_OFF_BAR + 256;            // the dynamic window starts 1 KB aligned (checked in the kernel)
constexpr uint32_t ATC_TMEM_COLS = 256;                // S buffers [0,192), O_h [192+16h, +16)
constexpr uint32_t ATC_O_COL = 192;
#ifndef ATC_POLY_HALF
#define ATC_POLY_HALF 1                                // 1: one more polynomial pair in every second 8-key group (POLY + 1/2 per four pairs)
#endif
constexpr int ATC_POLY_DEFAULT = 1;                    // exponentials per 4 pairs on the FMA pipe (SYNT_ATT_POLY overrides)
constexpr float ATC_LAZY = 8.0f;                       // rescale only when the max grows by more than 2^8
static_assert(2 * (ATC_SMEM + 1024) <= 228 * 1024, "two CTAs per SM must fit in shared memory");
// Measurement knock-outs (tools/ubench/att_knock.cu compiles this file with -DATC_KNOCK=<mask>; wrong results, timing only):
// 1 exponentials -> one multiply, 2 no P-tile stores, 4 no P.V MMAs, 8 no S MMAs, 16 no V^T builder stores, 32 no exact
// first-chunk maximum, 64 softmax warps do not wait for S tiles (and never rescale), 128 nor for free P tiles.  The product build has ATC_KNOCK == 0 and carries none of it.
#ifndef ATC_KNOCK
#define ATC_KNOCK 0
#endif
// nanosleep per retry in the waits of the roles that have ring slack (ns): A = TMA producer / V^T builder waiting for a free
// K/V stage (3 chunks = 12 units of slack), B = P.V issuer waiting for a P tile, C = S issuer waiting for a free S buffer
#ifndef ATC_REGS_ROLE
#define ATC_REGS_ROLE 48
#define ATC_REGS_SOFTMAX 96
#endif
#ifndef ATC_SLEEP_A
#define ATC_SLEEP_A 0
#endif
#ifndef ATC_SLEEP_B
#define ATC_SLEEP_B 0
#endif
#ifndef ATC_SLEEP_C
#define ATC_SLEEP_C 0
#endif

struct AttnTcMaps { CUtensorMap k; };

__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
// exp2 on the FMA pipe (Cody-Waite split + cubic minimax, relative error 7.6e-5, far below the bf16 rounding of P): x is split
// into its nearest integer n (low mantissa bits of x + 1.5*2^23) and f = x - n in [-0.5, 0.5]; 2^f by Horner; n is added to
// the exponent field with one integer multiply-add.  7 FMA-pipe instructions + 1 FMNMX against one 8-clk MUFU slot.
__device__ __forceinline__ float ex2_poly(float x) {
    x = fmaxf(x, -126.0f);                                   // below, the exponent arithmetic would wrap (P is ~0 there anyway)
    const float t = x + 12582912.0f;
    const float f = x - (t - 12582912.0f);
    float p = fmaf(f, 0.05520550534129143f, 0.24261397123336792f);
    p = fmaf(p, f, 0.6932547688484192f);
    p = fmaf(p, f, 0.9999276995658875f);
    return __int_as_float(__float_as_int(t) * 8388608 + __float_as_int(p));
}
// Measured (B200, round 2, final kernel with P in tensor memory): polynomial pairs per 8 pairs 0 / 2 / 3 / 4 -> 620 / 522 / 512 / 555 us
// per N=1024 launch, hence three in eight (POLY = 1 plus ATC_POLY_HALF).  Earlier (persistent kernel, scalar arithmetic):
// POLY 0 / 1 / 2 / 3 -> 672 / 653 / 727 / 942 us per launch: the
// softmax warps are balanced between the MUFU pipe and their issue slots, so one pair in four is the optimum.  A zero-reference
// fast path (P = 2^S without the subtraction while the row maximum stays within 2^+-40) measured SLOWER (714 us): the second
// copy of the unrolled loop costs more in registers / spills than the 64 FADDs per unit it removes.
// NOTE (measured, B200, round 1): moving 3 of every 8 exponentials to an FMA-pipe polynomial (Cody-Waite + degree-3
// minimax) made the kernel 10% SLOWER (4.09 -> 4.50 ms/step): the softmax warps are issue/FMA-pipe limited
// next to the MUFU pipe, and packed ex2.approx.{f16,bf16}x2 lowers to two MUFU ops on sm_100a.  All
// exponentials therefore stay on MUFU.EX2.
// Packed fp32x2 arithmetic (sm_100: add / mul / fma .f32x2 on 64-bit register pairs): one issue slot for two elements.  The
// softmax warps are bound by instruction issue next to the MUFU pipe (ncu: issue active 72%, XU 59%), so the subtraction of
// the reference and the whole polynomial run on pairs.
#ifndef ATC_PACKED
#define ATC_PACKED 1
#endif
// ex2_poly on a pair (same arithmetic per element as ex2_poly)
__device__ __forceinline__ float2 ex2_poly2(float2 x) {
    x.x = fmaxf(x.x, -126.0f); x.y = fmaxf(x.y, -126.0f);
    const float2 magic = make_float2(12582912.0f, 12582912.0f), nmagic = make_float2(-12582912.0f, -12582912.0f);
    const float2 t = add2(x, magic);
    const float2 n = add2(t, nmagic);
    const float2 f = add2(x, make_float2(-n.x, -n.y));
    float2 p = fma2(f, make_float2(0.05520550534129143f, 0.05520550534129143f), make_float2(0.24261397123336792f, 0.24261397123336792f));
    p = fma2(p, f, make_float2(0.6932547688484192f, 0.6932547688484192f));
    p = fma2(p, f, make_float2(0.9999276995658875f, 0.9999276995658875f));
    return make_float2(__int_as_float(__float_as_int(t.x) * 8388608 + __float_as_int(p.x)),
                       __int_as_float(__float_as_int(t.y) * 8388608 + __float_as_int(p.y)));
}
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
    uint32_t r;
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
    return r;
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_x16(uint32_t taddr, uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_st_x16(uint32_t taddr, const uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
        ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]),
          "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
        : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}

__device__ __forceinline__ void tmem_st_x8_nowait(uint32_t taddr, const uint32_t (&v)[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
                 ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]) : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// D[tmem] (+)= A[tmem] * B[smem]: M = 128 rows in the 128 lanes, K = 16 bf16 in 8 columns from tmem_a
__device__ __forceinline__ void umma_bf16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}

// PERSISTENT kernel: the grid is two CTAs per SM and every CTA walks the work items it, it + grid, it + 2*grid, ... (item =
// 128 queries x 4 heads of one image; all items cost the same).  Nothing is torn down between items: TMEM, the barriers
// and all the rings (K/V stages, S buffers, P tiles) run on, the TMA producer and the V^T builder prefetch across the item
// boundary, and only two hand-shakes are per item -- q_free / q_full around the rewrite of the query tile and o_full /
// o_free around the read-out of the accumulators.  (The one-CTA-per-item version lost 16% of the CTA slots' time between a
// CTA's exit and its successor's entry plus ~9% in every CTA's ramp)
struct AtcItem { int q0, hg, b; };
__device__ __forceinline__ AtcItem atc_item(int it, int nqt) {
    AtcItem o;
    o.q0 = (it % nqt) * 128;
    const int r = it / nqt;
    o.hg = r & 7;
    o.b = r >> 3;
    return o;
}

// POLY: of every four bf16 pairs of P, POLY are computed with ex2_poly on the FMA pipe instead of MUFU.EX2
template <int POLY, int HALF = 0>
__global__ void __launch_bounds__(ATC_THREADS, 2) attention_tc_kernel(const __grid_constant__ AttnTcMaps maps, int N,
                                                                      int C, int n_items, const bf16* __restrict__ qkv,
                                                                      bf16* __restrict__ out, int* __restrict__ work_ctr,
                                                                      long long* __restrict__ tl) {
    extern __shared__ __align__(1024) uint8_t smem[];
    if ((smem_u32(smem) & 1023u) != 0) __trap();             // SWIZZLE_128B tiles need 1 KB alignment
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + ATC_OFF_BAR);
    uint64_t* full_bar = bars;                   // [STAGES]
    uint64_t* empty_bar = bars + ATC_STAGES;     // [STAGES]
    uint64_t* s_full = bars + 2 * ATC_STAGES;    // [NS]
    uint64_t* s_free = s_full + ATC_NS;          // [NS]
    uint64_t* p_full = s_free + ATC_NS;          // [NS] P tiles rotate like the S buffers
    uint64_t* p_free = p_full + ATC_NS;          // [NS]
    uint64_t* q_full = p_free + ATC_NS;          // zero-masked Q tile of the item written by the softmax warps
    uint64_t* q_free = q_full + 1;               // every S MMA of the item has read the Q tile
    uint64_t* o_full = q_free + 1;               // every P.V MMA of the item is complete
    uint64_t* o_free = o_full + 1;               // the softmax warps have read the accumulators out
    uint64_t* item_full = o_free + 1;            // [4] work-queue ring: item id of the CTA's k-th item published
    int* item_ring = reinterpret_cast<int*>(item_full + 4);   // [4]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(item_ring + 4);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nqt = N / 128;
    const int n_chunks = N / ATC_KEYS;
    const int n_units = n_chunks * 4;            // unit u = chunk*4 + head; a multiple of 4, so head = global unit & 3 too
    // DYNAMIC work queue: the TMA producer thread draws item ids from a global counter (one ahead of the item it is loading)
    // and publishes them in a 4-deep shared-memory ring; every role reads its k-th item from slot k & 3.  The two CTAs of an
    // SM do NOT progress at the same rate (measured with a static round-robin: 61 k vs 131 k clk per item -- the warp
    // scheduler favours one CTA's softmax warps at the MUFU pipe), so a static split leaves the favoured CTA idle at the
    // end; with the queue it simply takes more items.  No role can be more than two items away from the producer (the K/V
    // stage ring is 3 chunks deep), so slot k & 3 is never overwritten while somebody still needs it.
    auto get_item = [&](int k) -> int {
        mbar_wait(&item_full[k & 3], ((uint32_t)k >> 2) & 1u);
        return *reinterpret_cast<volatile int*>(&item_ring[k & 3]);
    };

    pdl_launch_dependents();
    pdl_wait();                                              // qkv comes from the previous kernel
    // the softmax threads fetch their 32 bytes of the first item's query tile at once: the latency overlaps the set-up
    uint4 q_pre[2] = {make_uint4(0u, 0u, 0u, 0u), make_uint4(0u, 0u, 0u, 0u)};
    auto load_q = [&](const AtcItem& im) {
        const int t = threadIdx.x - 128, qrow = t >> 1, hp = t & 1;
        const uint4* src = reinterpret_cast<const uint4*>(qkv + ((size_t)im.b * N + im.q0 + qrow) * (3 * C) + im.hg * 32 + hp * 16);
        q_pre[0] = __ldg(src); q_pre[1] = __ldg(src + 1);
    };
    // (the first item of a CTA is only known after the queue is up: its query / v fetches follow the set-up barrier)
    // V^T builder (warp 3): constant rows and the first chunk's v values before the set-up barrier (latency overlap)
    uint4 nx[2][4];
    const size_t key_stride = (size_t)(3 * C) / 8;                            // uint4 per token row
    auto fetch = [&](const AtcItem& im, int c) {
        const uint4* vsrc = reinterpret_cast<const uint4*>(qkv + ((size_t)im.b * N + 2 * lane) * (3 * C) + 2 * C + im.hg * 32);
#pragma unroll
        for (int kk = 0; kk < 2; ++kk)
#pragma unroll
            for (int h = 0; h < 4; ++h) nx[kk][h] = __ldg(vsrc + ((size_t)c * ATC_KEYS + kk) * key_stride + h);
    };
    if (warp == 3) {
        for (int i = lane; i < ATC_STAGES * 4 * 8 * 32; i += 32) {            // rows 8..15 of every head of every slot
            const int word = i & 31, row = 8 + ((i >> 5) & 7), h = (i >> 8) & 3, st = i >> 10;
            uint8_t* base = smem + ATC_OFF_STAGE + st * ATC_STAGE_BYTES + ATC_K_BYTES + h * 2048 + row * 128;
            *reinterpret_cast<uint32_t*>(base + ((((word >> 2) ^ (row & 7)) << 4) | ((word & 3) << 2))) = row == 8 ? 0x3F803F80u : 0u;
        }
    }
    if (warp == 0 && lane == 0) {
        prefetch_tmap(&maps.k);
        // full: the TMA producer's arrive.expect_tx (K tile) + the V^T builder's arrive
        for (int s = 0; s < ATC_STAGES; ++s) { mbar_init(&full_bar[s], 2); mbar_init(&empty_bar[s], 1); }
        for (int s = 0; s < ATC_NS; ++s) { mbar_init(&s_full[s], 1); mbar_init(&s_free[s], ATC_P_TMEM ? 1 : 128); }
        for (int g = 0; g < ATC_NS; ++g) { mbar_init(&p_full[g], 128); mbar_init(&p_free[g], 1); }
        mbar_init(q_full, 256); mbar_init(q_free, 1); mbar_init(o_full, 1); mbar_init(o_free, 256);
        for (int i = 0; i < 4; ++i) mbar_init(&item_full[i], 1);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc<ATC_TMEM_COLS>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    // register re-allocation: the role warpgroup (warps 0..3) hands registers to the two softmax warpgroups, whose unrolled
    // unit body otherwise spills inside the hot loop (ncu: long-scoreboard stalls on the reloads)
    if (warp < 4) {
#if ATC_REGS_ROLE > 0
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(ATC_REGS_ROLE));
#endif
    if (warp == 0) {
        if (elect_one()) {
            // ===================== TMA producer =====================
            int stage = 0; uint32_t phase = 0;
            auto publish = [&](int k, int id) {
                *reinterpret_cast<volatile int*>(&item_ring[k & 3]) = id;
                mbar_arrive(&item_full[k & 3]);              // release: the id is visible to whoever sees the phase flip
            };
            int it = atomicAdd(work_ctr, 1);
            publish(0, it);
            for (int k = 0; it < n_items; ++k) {
                const int nxt = atomicAdd(work_ctr, 1);       // one ahead: the other roles prefetch across the item boundary
                publish(k + 1, nxt);
                const AtcItem im = atc_item(it, nqt);
                for (int c = 0; c < n_chunks; ++c) {
                    mbar_wait<ATC_SLEEP_A>(&empty_bar[stage], phase ^ 1u);
                    uint8_t* sk = smem + ATC_OFF_STAGE + stage * ATC_STAGE_BYTES;
                    mbar_arrive_expect_tx(&full_bar[stage], ATC_K_BYTES);
                    tma_load_2d(sk, &maps.k, &full_bar[stage], C + (im.hg >> 1) * 64, im.b * N + c * ATC_KEYS);
                    if (++stage == ATC_STAGES) { stage = 0; phase ^= 1u; }
                }
                it = nxt;
            }
        }
    } else if (warp == 3) {
        // ===================== V^T builder =====================
        // stage tile: [4 heads][16 rows][64 keys] bf16, SWIZZLE_128B (row = 128 B, 16-byte chunk ^= row & 7).  Lane l owns keys
        // 2l, 2l+1 of the chunk: their 4 x 8 v values arrive as 2 x 64 B from global memory and leave as 32 aligned 32-bit
        // stores (two adjacent keys of one row).  Rows 8 (ones) and 9..15 (zeros) are constant: written once per slot.
        int stage = 0; uint32_t phase = 0;
        int it = get_item(0);
        if (it < n_items) fetch(atc_item(it, nqt), 0);
        for (int k = 0; it < n_items; ++k) {
            const AtcItem im = atc_item(it, nqt);
            const int nxt = get_item(k + 1);
            for (int c = 0; c < n_chunks; ++c) {
                mbar_wait<ATC_SLEEP_A>(&empty_bar[stage], phase ^ 1u);
                uint8_t* sv = smem + ATC_OFF_STAGE + stage * ATC_STAGE_BYTES + ATC_K_BYTES;
#pragma unroll
                for (int h = 0; h < 4; ++h) {
                    const uint16_t* a = reinterpret_cast<const uint16_t*>(&nx[0][h]);     // key 2l:   dims 0..7 of head h
                    const uint16_t* bq = reinterpret_cast<const uint16_t*>(&nx[1][h]);    // key 2l+1
#pragma unroll
                    for (int d = 0; d < 8; ++d) {
                        const uint32_t pair = (uint32_t)a[d] | ((uint32_t)bq[d] << 16);
                        if (!(ATC_KNOCK & 16) || pair == 0x12345678u)
                        *reinterpret_cast<uint32_t*>(sv + h * 2048 + d * 128 + ((((lane >> 2) ^ d) << 4) | ((lane & 3) << 2))) = pair;
                    }
                }
                // the next chunk's values are requested after the stores (one register set; the ring keeps the builder up to
                // three chunks ahead of the MMAs, so the load latency hides behind the wait for the next free stage)
                if (c + 1 < n_chunks) fetch(im, c + 1);
                else if (nxt < n_items) fetch(atc_item(nxt, nqt), 0);                    // first chunk of the next item
                fence_proxy_async();                                              // generic-proxy writes -> UMMA (async proxy)
                __syncwarp();
                if (lane == 0) mbar_arrive(&full_bar[stage]);
                if (++stage == ATC_STAGES) { stage = 0; phase ^= 1u; }
            }
            it = nxt;
        }
    } else if (warp == 1) {
        if (elect_one()) {
            // ===================== S = Q K^T issuer =====================
            constexpr uint32_t idesc_s = make_idesc_bf16(128, ATC_KEYS);
            const uint32_t q_addr = smem_u32(smem);
            int stage = 0; uint32_t phase = 0;
            int sb = 0; uint32_t sph = 0;                    // S-buffer ring, runs across the items
            uint32_t iph = 0;                                // item parity
            for (int k = 0, it; (it = get_item(k)) < n_items; ++k) {
                const int hg = atc_item(it, nqt).hg;
                mbar_wait(q_full, iph);
                for (int u = 0; u < n_units; ++u) {
                    const int j = u & 3;
                    if (j == 0) { mbar_wait(&full_bar[stage], phase); }
                    mbar_wait<ATC_SLEEP_C>(&s_free[sb], sph ^ 1u);
                    tc_fence_after();
                    const uint32_t k_addr = smem_u32(smem + ATC_OFF_STAGE + stage * ATC_STAGE_BYTES);
                    // A: zero-masked head j of the Q tile; B: the natural 16-column pair that holds head j of this CTA's 4 heads
                    if (!(ATC_KNOCK & 8))
                    umma_bf16(tmem + sb * ATC_KEYS, make_smem_desc_sw128(q_addr) + 2 * j,
                              make_smem_desc_sw128(k_addr) + 4 * (hg & 1) + 2 * (j >> 1), idesc_s, 0u);
                    umma_commit(&s_full[sb]);
                    if (++sb == ATC_NS) { sb = 0; sph ^= 1u; }
                    if (j == 3) { if (++stage == ATC_STAGES) { stage = 0; phase ^= 1u; } }
                }
                umma_commit(q_free);                         // the query tile may be rewritten for the next item
                iph ^= 1u;
            }
        }
    } else if (warp == 2) {
        if (elect_one()) {
            // ===================== O += P V issuer =====================
            constexpr uint32_t idesc_pv = make_idesc_bf16(128, 16);
            const uint32_t p_addr = smem_u32(smem + ATC_OFF_P);
            int stage = 0; uint32_t phase = 0;
            int pb = 0; uint32_t pph = 0;                    // P-tile ring, runs across the items
            uint32_t iph = 0;
            for (int k = 0; get_item(k) < n_items; ++k) {
                if (k != 0) { mbar_wait(o_free, iph ^ 1u); tc_fence_after(); }   // previous item's accumulators were read out
                for (int u = 0; u < n_units; ++u) {
                    const int j = u & 3, c = u >> 2;
                    if (j == 0) mbar_wait(&full_bar[stage], phase);          // V of this chunk has landed
                    mbar_wait<ATC_SLEEP_B>(&p_full[pb], pph);
                    tc_fence_after();
                    const uint32_t v_addr = smem_u32(smem + ATC_OFF_STAGE + stage * ATC_STAGE_BYTES + ATC_K_BYTES);
                    const uint64_t dp = make_smem_desc_sw128(p_addr + pb * ATC_P_BYTES);
                    const uint64_t dv = make_smem_desc_sw128(v_addr + j * 2048);
#pragma unroll
                    for (int kk = 0; kk < 4; ++kk) {
                        if (ATC_KNOCK & 4) continue;
                        if (ATC_P_TMEM) umma_bf16_ts(tmem + ATC_O_COL + j * 16, tmem + pb * ATC_KEYS + 8 * kk, dv + 2 * kk, idesc_pv, (c | kk) != 0 ? 1u : 0u);
                        else umma_bf16(tmem + ATC_O_COL + j * 16, dp + 2 * kk, dv + 2 * kk, idesc_pv, (c | kk) != 0 ? 1u : 0u);
                    }
                    umma_commit(ATC_P_TMEM ? &s_free[pb] : &p_free[pb]);     // P in TMEM: this also frees the S buffer it lives in
                    if (++pb == ATC_NS) { pb = 0; pph ^= 1u; }
                    if (j == 3) {
                        umma_commit(&empty_bar[stage]);                      // all S and P.V reads of this stage are done
                        if (++stage == ATC_STAGES) { stage = 0; phase ^= 1u; }
                    }
                }
                umma_commit(o_full);
                iph ^= 1u;
            }
        }
    }
    } else {
#if ATC_REGS_ROLE > 0
        asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(ATC_REGS_SOFTMAX));
#endif
        // ===================== softmax warpgroups =====================
        const int g = (warp - 4) >> 2;                     // warpgroup 0 / 1
        const int quarter = warp & 3;                      // TMEM lane quarter of this warp
        const int r = quarter * 32 + lane;                 // query row
        const uint32_t lane_addr = tmem + ((uint32_t)(quarter * 32) << 16);
        const int sw = r & 7;
        // this warpgroup's units are g, g+2 (mod 4) of every chunk: ring position of unit (c*4 + g + 2*jj) advances by
        // 2 per unit handled here and by 4 per chunk; kept as (index mod 3, wrap parity) of the GLOBAL unit counter
        int sbu = g % ATC_NS; uint32_t sphu = 0;           // ring slot / parity of this warpgroup's next unit
        auto ring_advance2 = [&]() { sbu += 2; if (sbu >= ATC_NS) { sbu -= ATC_NS; sphu ^= 1u; } };
        uint32_t iph = 0;
#ifdef SYNT_ATT_TIMELINE_BUILD
        // debug build (tools/att_timeline.py): per CTA and softmax warp {smid, entry clock, then per item: end clock and the
        // cycles spent waiting for S tiles / free P tiles / the accumulators / the query tile}; 8 + 16*6 words per warp
        long long tw[4] = {0, 0, 0, 0}, tt0 = 0;
        int tl_item = 0;
        long long* trow = tl ? tl + ((size_t)blockIdx.x * 8 + (warp - 4)) * 128 : nullptr;
        if (trow && lane == 0) { unsigned smid; asm volatile("mov.u32 %0, %%smid;" : "=r"(smid)); trow[0] = smid; trow[1] = clock64(); }
#define ATC_T0() do { if (trow) tt0 = clock64(); } while (0)
#define ATC_T1(i) do { if (trow) tw[i] += clock64() - tt0; } while (0)
#define ATC_ITEM_END() do { if (trow && lane == 0 && tl_item < 20) { long long* o = trow + 8 + tl_item * 6; o[0] = clock64(); \
        o[1] = tw[0]; o[2] = tw[1]; o[3] = tw[2]; o[4] = tw[3]; o[5] = it; } ++tl_item; tw[0] = tw[1] = tw[2] = tw[3] = 0; } while (0)
#else
#define ATC_T0() do { } while (0)
#define ATC_T1(i) do { } while (0)
#define ATC_ITEM_END() do { } while (0)
#endif
        int it = get_item(0);
        if (it < n_items) load_q(atc_item(it, nqt));
        for (int k = 0; it < n_items; ++k) {
            const AtcItem im = atc_item(it, nqt);
            {
                // zero-masked Q tile: thread = (query row, pair of heads); head j keeps its 8 dims in chunk 2j + (j&1) of the
                // row (16-byte chunks, SWIZZLE_128B: physical chunk = logical ^ (row & 7)), the other chunk of the pair is 0
                const int t = threadIdx.x - 128, qrow = t >> 1, hp = t & 1;
                const uint4 h0 = q_pre[0], h1 = q_pre[1];                    // heads 2hp, 2hp+1
                ATC_T0();
                if (k != 0) mbar_wait(q_free, iph ^ 1u);                     // the previous item's S MMAs are done with the tile
                ATC_T1(3);
                uint8_t* qr = smem + qrow * 128;
                const int qs = qrow & 7;
                const uint4 z = make_uint4(0u, 0u, 0u, 0u);
                *reinterpret_cast<uint4*>(qr + (((4 * hp + 0) ^ qs) << 4)) = h0;    // head 2hp   -> (q | 0)
                *reinterpret_cast<uint4*>(qr + (((4 * hp + 1) ^ qs) << 4)) = z;
                *reinterpret_cast<uint4*>(qr + (((4 * hp + 2) ^ qs) << 4)) = z;     // head 2hp+1 -> (0 | q)
                *reinterpret_cast<uint4*>(qr + (((4 * hp + 3) ^ qs) << 4)) = h1;
                fence_proxy_async();
                mbar_arrive(q_full);
            }
            float m[2] = {0.f, 0.f};                           // running softmax reference of heads g and g+2
            float mnext[2] = {0.f, 0.f};                       // sampled maximum seen in the previous chunk
            for (int c = 0; c < n_chunks; ++c) {
#pragma unroll
                for (int jj = 0; jj < 2; ++jj) {
                    const int j = g + 2 * jj, sb = sbu;
                    const uint32_t sph = sphu;
                    ring_advance2();
                    ATC_T0();
                    if (!(ATC_KNOCK & 64)) mbar_wait(&s_full[sb], sph);
                    ATC_T1(0);
                    tc_fence_after();
                    // Softmax reference m of this row and head.  Any m within the bf16/fp32 exponent range of the true maximum
                    // gives the exact softmax after the final division: m starts at the maximum of the item's first 16 keys
                    // (below) and then follows a SAMPLED maximum (every 4th key) of the previous chunk, picked up while its
                    // exponentials were computed; it moves only when that exceeds m by more than 2^8 (lazy rescale).  P may
                    // therefore exceed 2^8 for a chunk; bf16 P / fp32 accumulators have the range.
                    float alpha = 1.0f;
                    bool moved = false;
                    if (c != 0 && !(ATC_KNOCK & 64) && mnext[jj] > m[jj] + ATC_LAZY) {
                        alpha = ex2_approx(m[jj] - mnext[jj]); m[jj] = mnext[jj]; moved = true;
                    }
                    uint8_t* p_row = smem + ATC_OFF_P + sb * ATC_P_BYTES + r * 128;
                    // P.V of the unit three before this one is complete, hence (in-order MMA pipe) so is every earlier one:
                    // P[sb] is free and O_j (last written four units ago) has no MMA in flight
                    ATC_T0();
                    // (P in TMEM: s_full of this unit already implies it -- the S MMA was issued behind the P.V commit that
                    // freed this buffer)
                    if (!ATC_P_TMEM && !(ATC_KNOCK & 128)) mbar_wait(&p_free[sb], sph ^ 1u);
                    ATC_T1(1);
                    if (__any_sync(0xffffffffu, moved)) {
                        tc_fence_after();
                        uint32_t o[16];
                        tmem_ld_x16(lane_addr + ATC_O_COL + j * 16, o);
                        tmem_ld_wait();
#pragma unroll
                        for (int i = 0; i < 16; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
                        tmem_st_x16(lane_addr + ATC_O_COL + j * 16, o);
                    }
                    // 16-column pieces, double-buffered in registers: the TMEM load of piece p+1 is in flight while the
                    // exponentials of piece p are computed (tcgen05.wait::ld waits for ALL loads, so it sits after the compute).
                    float smp;
                    {
                        uint32_t vv[2][16];
                        tmem_ld_x16(lane_addr + sb * ATC_KEYS, vv[0]);
                        tmem_ld_wait();
                        if (c == 0 && !(ATC_KNOCK & 32)) {
                            // first chunk of the item: the reference starts at the exact maximum of the chunk's first 16 keys
                            // (no second pass over the S tile -- a full-chunk maximum cost 5% of the kernel); from here on it
                            // follows the sampled maxima like in every later chunk
                            float t0 = fmaxf(fmaxf(__uint_as_float(vv[0][0]), __uint_as_float(vv[0][1])), __uint_as_float(vv[0][2]));
                            float t1 = fmaxf(fmaxf(__uint_as_float(vv[0][3]), __uint_as_float(vv[0][4])), __uint_as_float(vv[0][5]));
                            float t2 = fmaxf(fmaxf(__uint_as_float(vv[0][6]), __uint_as_float(vv[0][7])), __uint_as_float(vv[0][8]));
                            float t3 = fmaxf(fmaxf(__uint_as_float(vv[0][9]), __uint_as_float(vv[0][10])), __uint_as_float(vv[0][11]));
                            t0 = fmaxf(fmaxf(t0, __uint_as_float(vv[0][12])), __uint_as_float(vv[0][13]));
                            t1 = fmaxf(fmaxf(t1, __uint_as_float(vv[0][14])), __uint_as_float(vv[0][15]));
                            m[jj] = fmaxf(fmaxf(t0, t1), fmaxf(t2, t3));
                        }
                        const float mrow = m[jj];
                        smp = mrow;                                            // sampled maximum of this chunk (raw logits)
#pragma unroll
                        for (int piece = 0; piece < ATC_KEYS / 16; ++piece) {
                            uint32_t (&cur)[16] = vv[piece & 1];
                            if (piece + 1 < ATC_KEYS / 16) tmem_ld_x16(lane_addr + sb * ATC_KEYS + (piece + 1) * 16, vv[(piece + 1) & 1]);
                            smp = fmaxf(fmaxf(smp, __uint_as_float(cur[0])), __uint_as_float(cur[4]));
                            smp = fmaxf(fmaxf(smp, __uint_as_float(cur[8])), __uint_as_float(cur[12]));
#pragma unroll
                            uint32_t pw[8];                                        // the piece's 16 probabilities, packed
#pragma unroll
                            for (int q = 0; q < 2; ++q) {                          // 16-byte chunk = 8 keys
                                uint32_t (&w)[4] = *reinterpret_cast<uint32_t (*)[4]>(&pw[q * 4]);
#pragma unroll
                                for (int i = 0; i < 4; ++i) {
                                    const int e = q * 8 + i * 2;
#if ATC_PACKED
                                    const float2 xx = add2(make_float2(__uint_as_float(cur[e]), __uint_as_float(cur[e + 1])), make_float2(-mrow, -mrow));
                                    const float x0 = xx.x, x1 = xx.y;
                                    if (!(ATC_KNOCK & 1) && i >= 4 - POLY - (HALF && q == 1 ? 1 : 0)) { const float2 y = ex2_poly2(xx); w[i] = pack_bf16x2(y.x, y.y); continue; }
#else
                                    const float x0 = __uint_as_float(cur[e]) - mrow, x1 = __uint_as_float(cur[e + 1]) - mrow;
#endif
                                    if (ATC_KNOCK & 1) w[i] = pack_bf16x2(x0 * 0.5f, x1 * 0.5f);
                                    else
                                    w[i] = (POLY > 0 && i >= 4 - POLY) ? pack_bf16x2(ex2_poly(x0), ex2_poly(x1))
                                                                        : pack_bf16x2(ex2_approx(x0), ex2_approx(x1));
                                }
                                if (!ATC_P_TMEM && (!(ATC_KNOCK & 2) || w[0] == 0x12345678u))
                                *reinterpret_cast<uint4*>(p_row + (((piece * 2 + q) ^ sw) << 4)) = make_uint4(w[0], w[1], w[2], w[3]);   // SWIZZLE_128B
                            }
                            // P in TMEM: the 16 keys of this piece -> columns 8 piece .. 8 piece + 7 of the S buffer; they were S
                            // columns of a piece <= this one, i.e. already in registers (the load in flight reads piece + 1)
                            if (ATC_P_TMEM && !(ATC_KNOCK & 2)) tmem_st_x8_nowait(lane_addr + sb * ATC_KEYS + piece * 8, pw);
                            if (piece + 1 < ATC_KEYS / 16) {
                                tmem_ld_wait();
                                if (!ATC_P_TMEM && piece + 2 == ATC_KEYS / 16) {   // the last load of this S buffer has completed
                                    tc_fence_before();
                                    mbar_arrive(&s_free[sb]);
                                }
                            }
                        }
                    }
                    mnext[jj] = smp;
                    if (ATC_P_TMEM) tmem_st_wait();
                    tc_fence_before();
                    if (!ATC_P_TMEM) fence_proxy_async();                      // generic-proxy writes -> async proxy (UMMA)
                    mbar_arrive(&p_full[sb]);
                }
            }
            // ---- item boundary: the next item's queries are requested before the read-out so that their latency overlaps it
            const int nxt = get_item(k + 1);
            if (nxt < n_items) load_q(atc_item(nxt, nqt));
            // ---- epilogue: O_h / rowsum -> bf16 NHWC
            ATC_T0();
            mbar_wait(o_full, iph);
            ATC_T1(2);
            tc_fence_after();
            uint32_t ov[2][16];
#pragma unroll
            for (int jj = 0; jj < 2; ++jj) tmem_ld_x16(lane_addr + ATC_O_COL + (g + 2 * jj) * 16, ov[jj]);
            tmem_ld_wait();
            tc_fence_before();
            mbar_arrive(o_free);                               // the P.V issuer may start the next item's accumulation
#pragma unroll
            for (int jj = 0; jj < 2; ++jj) {
                const int j = g + 2 * jj;
                const float inv = 1.0f / __uint_as_float(ov[jj][8]);
                float o[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) o[i] = __uint_as_float(ov[jj][i]) * inv;
                store8<bf16>(out + ((size_t)im.b * N + im.q0 + r) * C + (im.hg * 4 + j) * 8, o);
            }
            iph ^= 1u;
            ATC_ITEM_END();
            it = nxt;
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc<ATC_TMEM_COLS>(tmem);
    // the last CTA to leave re-arms the queue for the next launch that uses this counter pair (every draw of every CTA
    // precedes that CTA's arrival here)
    if (threadIdx.x == 0) {
        __threadfence();
        if (atomicAdd(work_ctr + 1, 1) == (int)gridDim.x - 1) { atomicExch(work_ctr, 0); atomicExch(work_ctr + 1, 0); }
    }
}

// ---- V^T builder: vt[b*H + h][16][N] from the v part of qkv' -------------------------------
__global__ void __launch_bounds__(256) build_vt_kernel(const bf16* __restrict__ qkv, int N, int C, int ldq, int voff,
                                                       bf16* __restrict__ vt) {
    __shared__ bf16 tile[64][256 + 8];
    pdl_enter();
    const int b = blockIdx.y, t0 = blockIdx.x * 64;
    const int H = C / 8;
    for (int i = threadIdx.x; i < 64 * (C / 8); i += 256) {
        const int tok = i / (C / 8), v8 = i % (C / 8);
        *reinterpret_cast<uint4*>(&tile[tok][v8 * 8]) =
            *reinterpret_cast<const uint4*>(qkv + ((size_t)b * N + t0 + tok) * ldq + voff + v8 * 8);
    }
    __syncthreads();
    // each thread writes 8 consecutive tokens (16 B) of one (head, row)
    for (int i = threadIdx.x; i < H * 16 * 8; i += 256) {
        const int seg = i & 7, row = (i >> 3) & 15, h = i >> 7;
        __align__(16) bf16 vals[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            bf16 x;
            if (row < 8) x = tile[seg * 8 + k][h * 8 + row];
            else x = __float2bfloat16_rn(row == 8 ? 1.0f : 0.0f);
            vals[k] = x;
        }
        *reinterpret_cast<uint4*>(vt + (((size_t)b * H + h) * 16 + row) * N + t0 + seg * 8) = *reinterpret_cast<const uint4*>(vals);
    }
}

bool attention_tc_supported(int N, int C) { return (N % 128 == 0) && C == 256; }

// qkv: [B, N, 768] bf16 (q | k | v, q pre-scaled), vt scratch: [B*32, 16, N] bf16, out: [B, N, 256] bf16
void attention_tc(const void* qkv, int B, int N, int C, void* vt_scratch, void* out, cudaStream_t s) {
    SYNT_CHECK(attention_tc_supported(N, C), "attention_tc: unsupported shape");
    const int ldq = 3 * C;
    (void)vt_scratch;                                               // V^T is built inside the kernel
    AttnTcMaps maps;
    {
        cuuint64_t dims[2] = {(cuuint64_t)ldq, (cuuint64_t)B * N};
        cuuint64_t strides[1] = {(cuuint64_t)ldq * 2};
        cuuint32_t boxk[2] = {64, ATC_KEYS};
        encode_bf16_sw128(&maps.k, qkv, 2, dims, strides, boxk, "attention k");
    }
    static const int poly = [] { const char* e = getenv("SYNT_ATT_POLY"); const int v = e ? atoi(e) : ATC_POLY_DEFAULT; return v < 0 ? 0 : (v > 3 ? 3 : v); }();
    static const int half = [] { const char* e = getenv("SYNT_ATT_POLY_HALF"); return e ? (atoi(e) != 0) : (ATC_POLY_HALF != 0); }();
    auto kern = poly == 0 ? attention_tc_kernel<0> : poly == 1 ? (half ? attention_tc_kernel<1, 1> : attention_tc_kernel<1>)
              : poly == 2 ? attention_tc_kernel<2> : attention_tc_kernel<3>;
    ensure_dynamic_smem((const void*)(kern), ATC_SMEM);
    static const int num_sms = [] { int d = 0, n = 0; cudaGetDevice(&d); cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, d); return n; }();
    const int n_items = (N / 128) * (C / 32) * B;            // 128 queries x 4 heads of one image each
    const int grid = n_items < 2 * num_sms ? n_items : 2 * num_sms;     // persistent: two resident CTAs per SM
    // work-queue counters {next item, CTAs done}: a pool of pairs handed out round-robin, zero at rest (the kernel's last CTA
    // re-arms its pair), so that launches captured into different graphs / running on different streams do not share one
    constexpr int kCtrPairs = 1024, kMaxDev = 64;
    static int* ctr_pools[kMaxDev] = {nullptr};                      // one pool per device (a process may drive several GPUs)
    static int ctr_next[kMaxDev] = {0};
    static std::mutex ctr_mu;
    int dev = 0;
    SYNT_CUDA(cudaGetDevice(&dev));
    SYNT_CHECK(dev >= 0 && dev < kMaxDev, "attention_tc: device index");
    int* work_ctr;
    {
        std::lock_guard<std::mutex> lk(ctr_mu);
        if (!ctr_pools[dev]) {
            SYNT_CUDA(cudaMalloc(&ctr_pools[dev], kCtrPairs * 2 * sizeof(int)));
            SYNT_CUDA(cudaMemset(ctr_pools[dev], 0, kCtrPairs * 2 * sizeof(int)));
        }
        work_ctr = ctr_pools[dev] + 2 * (ctr_next[dev]++ % kCtrPairs);
    }
    long long* tl = nullptr;
#ifdef SYNT_ATT_TIMELINE_BUILD
    static const char* tl_path = getenv("SYNT_ATT_TIMELINE");
    const size_t tl_n = (size_t)grid * 8 * 128;
    if (tl_path && N == 1024 && B > 1) { SYNT_CUDA(cudaMalloc(&tl, tl_n * 8)); SYNT_CUDA(cudaMemsetAsync(tl, 0, tl_n * 8, s)); }
#endif
    launch_pdl(kern, dim3(grid), dim3(ATC_THREADS), ATC_SMEM, s, maps, N, C, n_items, (const bf16*)qkv, (bf16*)out,
               work_ctr, tl);
#ifdef SYNT_ATT_TIMELINE_BUILD
    if (tl) {
        std::vector<long long> h(tl_n);
        SYNT_CUDA(cudaMemcpyAsync(h.data(), tl, tl_n * 8, cudaMemcpyDeviceToHost, s));
        SYNT_CUDA(cudaStreamSynchronize(s));
        cudaFree(tl);
        if (FILE* f = fopen(tl_path, "wb")) { fwrite(h.data(), 8, tl_n, f); fclose(f); }
    }
#endif
}

}  // namespace synt
