// Region selection on attribution maps (select_regions_advanced, xai/XAI.py:1340-1451) -- SURVEY.md section 8f row 4.
// The reference does this per map on the CPU with numpy / scipy.ndimage: channel L2 norm -> np.percentile threshold ->
// binary closing (2 iterations) -> binary opening -> connected-component labelling -> drop components smaller than
// max(10, 1% of the pixels) -> statistics.  Here ONE CTA handles one map entirely in shared memory, any number of maps per
// launch (integer / comparison work, bit-exact masks):
//
//   saliency    s = sqrt((x0^2 + x1^2) + x2^2) with individually rounded fp32 operations (numpy's reduction order), or |x|
//   threshold   bitonic sort of the <= 16384 saliencies, then numpy's `linear` percentile EXACTLY as numpy >= 2 evaluates it
//               for a float32 array: virtual index (n-1)*q and the interpolation weight in float32, _lerp with the
//               t >= 0.5 form (every operation rounded separately: no FMA contraction)
//   morphology  dilate, dilate, erode, erode (closing x2), erode, dilate (opening), 3x3 square (connectivity 8) or cross
//               (connectivity 4); pixels outside the image are background for both operations (scipy's border_value=0)
//   components  min-label propagation with pointer jumping on the shared-memory label image (labels only decrease, so
//               the racy in-place sweep converges to the smallest pixel index of each component), sizes by shared atomics
//   statistics  count, mean / std of the saliency, mean / std / max / min over the selection (fp64 accumulation)
#include "kernels.cuh"
#include "../../include/synt_isic.h"
#include <string>

namespace synt {

extern thread_local std::string g_last_error;

constexpr int RG_THREADS = 1024, RG_MAX = 16384;
constexpr int RG_SMEM = RG_MAX * 4 /*sort keys -> labels*/ + (RG_MAX + 1) * 4 /*component sizes*/ + 2 * RG_MAX /*masks*/ + 64;

__device__ __forceinline__ float rg_saliency(const float* __restrict__ a, int C, int n, int i, int use_abs) {
    if (use_abs) return fabsf(a[i]);
    float s = __fmul_rn(a[i], a[i]);
    for (int c = 1; c < C; ++c) s = __fadd_rn(s, __fmul_rn(a[(size_t)c * n + i], a[(size_t)c * n + i]));
    return __fsqrt_rn(s);
}

__device__ __forceinline__ double rg_block_sum(double v, double* red) {
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    double t = 0.0;
    for (int w = 0; w < RG_THREADS / 32; ++w) t += red[w];
    return t;
}

__global__ void __launch_bounds__(RG_THREADS, 1) select_regions_kernel(const float* __restrict__ attr, int C, int H, int W,
                                                                       int use_abs, double q_percent, int bottom, int morphology,
                                                                       int conn4, unsigned char* __restrict__ mask_out,
                                                                       double* __restrict__ stats_out) {
    extern __shared__ __align__(16) unsigned char rg_smem[];
    float* keys = reinterpret_cast<float*>(rg_smem);
    int* lbl = reinterpret_cast<int*>(rg_smem);                                 // reuses the sort buffer
    int* cnt = reinterpret_cast<int*>(rg_smem + RG_MAX * 4);
    unsigned char* m0 = rg_smem + RG_MAX * 4 + (RG_MAX + 1) * 4 + 12;          // 16-byte aligned: 65536 + 65540 + 12
    unsigned char* m1 = m0 + RG_MAX;
    __shared__ float thr_s;
    __shared__ double red[RG_THREADS / 32];
    const int n = H * W, tid = threadIdx.x;
    const float* a = attr + (size_t)blockIdx.x * (use_abs ? 1 : C) * n;
    int n2 = 1;
    while (n2 < n) n2 <<= 1;

    // ---- threshold: sort, then numpy's float32 `linear` percentile
    for (int i = tid; i < n2; i += RG_THREADS) keys[i] = i < n ? rg_saliency(a, C, n, i, use_abs) : INFINITY;
    __syncthreads();
    for (int k = 2; k <= n2; k <<= 1)
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = tid; i < n2; i += RG_THREADS) {
                const int p = i ^ j;
                if (p > i) {
                    const float x = keys[i], y = keys[p];
                    const bool asc = (i & k) == 0;
                    if ((x > y) == asc) { keys[i] = y; keys[p] = x; }
                }
            }
            __syncthreads();
        }
    if (tid == 0) {
        const float q32 = (float)(q_percent / 100.0);
        const float v = __fmul_rn((float)(n - 1), q32);
        int lo = (int)floorf(v);
        lo = lo < 0 ? 0 : (lo > n - 1 ? n - 1 : lo);
        const int hi = lo + 1 > n - 1 ? n - 1 : lo + 1;
        const float t = __fsub_rn(v, (float)lo);
        const float A = keys[lo], B = keys[hi];
        const float d = __fsub_rn(B, A);
        float y = __fadd_rn(A, __fmul_rn(d, t));
        if (t >= 0.5f) y = __fsub_rn(B, __fmul_rn(d, __fsub_rn(1.0f, t)));
        thr_s = y;
    }
    __syncthreads();
    const float thr = thr_s;
    __syncthreads();                                                             // keys are dead from here on (lbl aliases them)

    unsigned char* src = m0;
    unsigned char* dst = m1;
    for (int i = tid; i < n; i += RG_THREADS) {
        const float s = rg_saliency(a, C, n, i, use_abs);
        src[i] = bottom ? (s <= thr) : (s >= thr);
    }
    __syncthreads();

    if (morphology) {
        // ---- closing x2 (D D E E), opening (E D); outside of the image = background
        const int ops[6] = {1, 1, 0, 0, 0, 1};                                  // 1 dilate, 0 erode
        for (int pass = 0; pass < 6; ++pass) {
            const int dil = ops[pass];
            for (int i = tid; i < n; i += RG_THREADS) {
                const int y = i / W, x = i - y * W;
                int acc = dil ? 0 : 1;
                for (int dy = -1; dy <= 1; ++dy)
                    for (int dx = -1; dx <= 1; ++dx) {
                        if (conn4 && dy != 0 && dx != 0) continue;
                        const int yy = y + dy, xx = x + dx;
                        const int v = (yy >= 0 && yy < H && xx >= 0 && xx < W) ? src[yy * W + xx] : 0;
                        acc = dil ? (acc | v) : (acc & v);
                    }
                dst[i] = (unsigned char)acc;
            }
            __syncthreads();
            unsigned char* tmp = src; src = dst; dst = tmp;
        }
        // ---- connected components (same structuring element), sizes, small-component removal
        for (int i = tid; i < n; i += RG_THREADS) { lbl[i] = src[i] ? i : n; cnt[i] = 0; }
        if (tid == 0) cnt[n] = 0;
        __syncthreads();
        int changed;
        do {
            changed = 0;
            for (int i = tid; i < n; i += RG_THREADS) {
                if (!src[i]) continue;
                const int y = i / W, x = i - y * W;
                const int cur = *(volatile int*)(lbl + i);
                int m = cur;
                for (int dy = -1; dy <= 1; ++dy)
                    for (int dx = -1; dx <= 1; ++dx) {
                        if ((conn4 && dy != 0 && dx != 0) || (dy == 0 && dx == 0)) continue;
                        const int yy = y + dy, xx = x + dx;
                        if (yy < 0 || yy >= H || xx < 0 || xx >= W || !src[yy * W + xx]) continue;
                        const int l = *(volatile int*)(lbl + yy * W + xx);
                        m = l < m ? l : m;
                    }
                const int r = *(volatile int*)(lbl + m);                         // pointer jump (label of the label)
                m = r < m ? r : m;
                if (m < cur) { atomicMin(lbl + i, m); changed = 1; }
            }
            changed = __syncthreads_or(changed);
        } while (changed);
        for (int i = tid; i < n; i += RG_THREADS)
            if (src[i]) atomicAdd(cnt + lbl[i], 1);
        __syncthreads();
        int min_size = (int)(0.01 * (double)n);
        if (min_size < 10) min_size = 10;
        for (int i = tid; i < n; i += RG_THREADS) dst[i] = (src[i] && cnt[lbl[i]] >= min_size) ? 1 : 0;
        __syncthreads();
        src = dst;
    }

    // ---- output + statistics
    double s1 = 0.0, s2 = 0.0, c = 0.0, t1 = 0.0, t2 = 0.0;
    float mx = -INFINITY, mn = INFINITY;
    for (int i = tid; i < n; i += RG_THREADS) {
        const float s = rg_saliency(a, C, n, i, use_abs);
        const int sel = src[i];
        mask_out[(size_t)blockIdx.x * n + i] = (unsigned char)sel;
        s1 += s; s2 += (double)s * s;
        if (sel) { c += 1.0; t1 += s; t2 += (double)s * s; mx = fmaxf(mx, s); mn = fminf(mn, s); }
    }
    s1 = rg_block_sum(s1, red); s2 = rg_block_sum(s2, red); c = rg_block_sum(c, red);
    t1 = rg_block_sum(t1, red); t2 = rg_block_sum(t2, red);
    for (int o = 16; o > 0; o >>= 1) {
        mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
    }
    __shared__ float mxs[RG_THREADS / 32], mns[RG_THREADS / 32];
    __syncthreads();
    if ((tid & 31) == 0) { mxs[tid >> 5] = mx; mns[tid >> 5] = mn; }
    __syncthreads();
    if (tid == 0) {
        for (int w = 0; w < RG_THREADS / 32; ++w) { mx = fmaxf(mx, mxs[w]); mn = fminf(mn, mns[w]); }
        double* o = stats_out + (size_t)blockIdx.x * 8;
        const double mean = s1 / n, var = s2 / n - mean * mean;
        o[0] = c; o[1] = (double)thr; o[2] = mean; o[3] = sqrt(var > 0.0 ? var : 0.0);
        if (c > 0.0) {
            const double ms = t1 / c, vs = t2 / c - ms * ms;
            o[4] = ms; o[5] = sqrt(vs > 0.0 ? vs : 0.0); o[6] = (double)mx; o[7] = (double)mn;
        } else {
            o[4] = o[5] = o[6] = o[7] = 0.0;
        }
    }
}

void select_regions(const float* attr, int n_maps, int C, int H, int W, int use_abs, double q_percent, int bottom, int morphology,
                    int connectivity, unsigned char* mask, double* stats, cudaStream_t s) {
    SYNT_CHECK(H > 0 && W > 0 && H * W <= RG_MAX, "select_regions: maps of at most 16384 pixels");
    SYNT_CHECK(C >= 1, "select_regions: C >= 1");
    SYNT_CHECK(connectivity == 4 || connectivity == 8, "select_regions: connectivity 4 or 8");
    SYNT_CHECK(q_percent >= 0.0 && q_percent <= 100.0, "select_regions: percentile outside [0, 100]");
    ensure_dynamic_smem((const void*)(select_regions_kernel), RG_SMEM);
    select_regions_kernel<<<n_maps, RG_THREADS, RG_SMEM, s>>>(attr, C, H, W, use_abs, q_percent, bottom, morphology,
                                                              connectivity == 4, mask, stats);
    SYNT_LAUNCH_CHECK();
}

}  // namespace synt

extern "C" int synt_select_regions(const float* attr_dev, int n_maps, int C, int H, int W, int use_abs, double k_percent,
                                   int bottom, int morphology, int connectivity, unsigned char* mask_dev, double* stats_dev,
                                   void* stream) {
    try {
        SYNT_CHECK(attr_dev && mask_dev && stats_dev && n_maps > 0, "bad argument");
        synt::select_regions(attr_dev, n_maps, C, H, W, use_abs, bottom ? k_percent : 100.0 - k_percent, bottom, morphology,
                             connectivity, mask_dev, stats_dev, (cudaStream_t)stream);
    } catch (const synt::Error& e) { synt::g_last_error = e.what(); return e.code;
    } catch (const std::exception& e) { synt::g_last_error = e.what(); return -1; }
    return 0;
}
