// Resampling loops of the reference's statistical validation (xai/XAI.py:1708-2005, statistical_validation_comprehensive):
//   bootstrap   (:1852-1878)  1000 x { resample both CFI samples with replacement, difference of the means }
//   permutation (:1882-1913)  10000 x { shuffle the pooled sample, difference of the means of the two parts }
// The reference runs them as Python loops over numpy calls; here every replicate is one thread.  The means are float64 with
// numpy's own summation order (pairwise_sum: < 8 sequential, <= 128 eight interleaved accumulators, above that recursive
// halves), so that with INJECTED resampling indices / permutations the replicate differences are bit-identical to the
// reference's; without them the indices come from an in-kernel Philox stream (statistically equivalent, not numpy's MT19937
// stream).  The closed-form tests on the few dozen scalars (t, Welch, Mann-Whitney, ...) stay scipy calls on the host.
#include "kernels.cuh"
#include "../../include/synt_isic.h"

namespace synt {

extern thread_local std::string g_last_error;

// numpy DOUBLE_pairwise_sum over a(i), i in [lo, lo + n): blocks of <= 128 values are summed with eight interleaved
// accumulators, larger ranges are split in halves (first half rounded down to a multiple of 8) and the two sums added --
// evaluated here with an explicit stack instead of recursion (a recursive device function would run on the 1 KB default
// CUDA stack next to the 2 KB shuffle array).
template <typename F>
__device__ double np_block_sum(F a, int lo, int n) {
    if (n < 8) {
        double r = 0.0;
        for (int i = 0; i < n; ++i) r += a(lo + i);
        return r;
    }
    double r[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) r[j] = a(lo + j);
    int i = 8;
    for (; i < n - (n % 8); i += 8) {
#pragma unroll
        for (int j = 0; j < 8; ++j) r[j] += a(lo + i + j);
    }
    double res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
    for (; i < n; ++i) res += a(lo + i);
    return res;
}
template <typename F>
__device__ double np_pairwise_sum(F a, int lo, int n) {
    if (n <= 128) return np_block_sum(a, lo, n);
    // post-order walk of the split tree: frame = (lo, n, value of the left child once known)
    constexpr int DEPTH = 24;
    int f_lo[DEPTH], f_n[DEPTH], f_state[DEPTH];
    double f_left[DEPTH];
    int sp = 0;
    f_lo[0] = lo; f_n[0] = n; f_state[0] = 0; f_left[0] = 0.0;
    double ret = 0.0;
    while (sp >= 0) {
        const int cn = f_n[sp], clo = f_lo[sp];
        if (cn <= 128) { ret = np_block_sum(a, clo, cn); --sp; continue; }
        int n2 = cn / 2;
        n2 -= n2 % 8;
        if (f_state[sp] == 0) {                              // descend into the left half
            f_state[sp] = 1;
            ++sp; f_lo[sp] = clo; f_n[sp] = n2; f_state[sp] = 0;
        } else if (f_state[sp] == 1) {                       // left done: keep it, descend into the right half
            f_left[sp] = ret; f_state[sp] = 2;
            ++sp; f_lo[sp] = clo + n2; f_n[sp] = cn - n2; f_state[sp] = 0;
        } else {                                             // both done
            ret = f_left[sp] + ret; --sp;
        }
    }
    return ret;
}

__device__ __forceinline__ uint32_t philox_u32(unsigned long long seed, unsigned long long stream, uint32_t ctr) {
    uint32_t c[4] = {ctr, (uint32_t)stream, (uint32_t)(stream >> 32), 0x57a75u};
    uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c[0]), lo0 = 0xD2511F53u * c[0];
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c[2]), lo1 = 0xCD9E8D57u * c[2];
        c[0] = hi1 ^ c[1] ^ k0; c[1] = lo1; c[2] = hi0 ^ c[3] ^ k1; c[3] = lo0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    return c[0];
}

constexpr int ST_MAX_N = 1024;                              // pooled sample size limit of the in-kernel shuffle

// out[b] = mean(top[it[b][:]]) - mean(bottom[ib[b][:]])
__global__ void bootstrap_kernel(const double* __restrict__ top, int n1, const double* __restrict__ bottom, int n2,
                                 const int* __restrict__ idx_top, const int* __restrict__ idx_bot, unsigned long long seed,
                                 int n_rep, double* __restrict__ out) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= n_rep) return;
    auto at = [&](int i) {
        const int k = idx_top ? idx_top[(size_t)b * n1 + i] : (int)(((unsigned long long)philox_u32(seed, 2ull * b, i) * (unsigned)n1) >> 32);
        return top[k];
    };
    auto ab = [&](int i) {
        const int k = idx_bot ? idx_bot[(size_t)b * n2 + i] : (int)(((unsigned long long)philox_u32(seed, 2ull * b + 1, i) * (unsigned)n2) >> 32);
        return bottom[k];
    };
    out[b] = np_pairwise_sum(at, 0, n1) / (double)n1 - np_pairwise_sum(ab, 0, n2) / (double)n2;
}

// out[b] = mean(combined[perm[b][:n1]]) - mean(combined[perm[b][n1:]])
__global__ void permutation_kernel(const double* __restrict__ combined, int n, int n1, const int* __restrict__ perms,
                                   unsigned long long seed, int n_rep, double* __restrict__ out) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= n_rep) return;
    unsigned short p[ST_MAX_N];
    if (!perms) {                                            // Fisher-Yates from the Philox stream of this replicate
        for (int i = 0; i < n; ++i) p[i] = (unsigned short)i;
        for (int i = n - 1; i > 0; --i) {
            const int j = (int)(((unsigned long long)philox_u32(seed, b, i) * (unsigned)(i + 1)) >> 32);
            const unsigned short t = p[i]; p[i] = p[j]; p[j] = t;
        }
    }
    auto a = [&](int i) { return combined[perms ? perms[(size_t)b * n + i] : p[i]]; };
    out[b] = np_pairwise_sum(a, 0, n1) / (double)n1 - np_pairwise_sum(a, n1, n - n1) / (double)(n - n1);
}

}  // namespace synt

using namespace synt;

extern "C" {

int synt_stat_bootstrap_mean_diff(const double* top_dev, int n1, const double* bottom_dev, int n2, const int* idx_top_dev,
                                  const int* idx_bottom_dev, unsigned long long seed, int n_bootstrap, double* out_dev,
                                  void* stream) {
    try {
        SYNT_CHECK(top_dev && bottom_dev && out_dev && n1 > 0 && n2 > 0 && n_bootstrap > 0, "bad argument");
        SYNT_CHECK((idx_top_dev == nullptr) == (idx_bottom_dev == nullptr), "inject both index arrays or none");
        bootstrap_kernel<<<(n_bootstrap + 127) / 128, 128, 0, (cudaStream_t)stream>>>(top_dev, n1, bottom_dev, n2, idx_top_dev,
                                                                                      idx_bottom_dev, seed, n_bootstrap, out_dev);
        SYNT_LAUNCH_CHECK();
    } catch (const synt::Error& e) { synt::g_last_error = e.what(); return e.code;
    } catch (const std::exception& e) { synt::g_last_error = e.what(); return -1; }
    return 0;
}

int synt_stat_permutation_mean_diff(const double* combined_dev, int n, int n1, const int* perms_dev, unsigned long long seed,
                                    int n_permutations, double* out_dev, void* stream) {
    try {
        SYNT_CHECK(combined_dev && out_dev && n1 > 0 && n > n1 && n_permutations > 0, "bad argument");
        SYNT_CHECK(perms_dev != nullptr || n <= ST_MAX_N, "in-kernel shuffle: pooled sample limited to 1024 values");
        permutation_kernel<<<(n_permutations + 127) / 128, 128, 0, (cudaStream_t)stream>>>(combined_dev, n, n1, perms_dev, seed,
                                                                                           n_permutations, out_dev);
        SYNT_LAUNCH_CHECK();
    } catch (const synt::Error& e) { synt::g_last_error = e.what(); return e.code;
    } catch (const std::exception& e) { synt::g_last_error = e.what(); return -1; }
    return 0;
}

}  // extern "C"
