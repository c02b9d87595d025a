// Shared helpers for the synt_isic_b200 CUDA library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cstdint>
#include <cstdio>
#include <string>
#include <stdexcept>
#include <mutex>
#include <set>
#include <utility>

namespace synt {

// ---- error plumbing: C++ exceptions inside, int codes at the C-ABI ------------------
struct Error : std::runtime_error {
    int code;
    Error(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};

#define SYNT_CUDA(expr)                                                                  \
    do {                                                                                 \
        cudaError_t _e = (expr);                                                         \
        if (_e != cudaSuccess)                                                           \
            throw ::synt::Error(-2, std::string(#expr) + " -> " + cudaGetErrorString(_e) + \
                                        " at " + __FILE__ + ":" + std::to_string(__LINE__)); \
    } while (0)

#define SYNT_CHECK(cond, msg)                                                            \
    do {                                                                                 \
        if (!(cond))                                                                     \
            throw ::synt::Error(-1, std::string(msg) + " (" #cond ") at " + __FILE__ + ":" + \
                                        std::to_string(__LINE__));                       \
    } while (0)

#define SYNT_LAUNCH_CHECK() SYNT_CUDA(cudaGetLastError())

// ---- storage-type helpers ------------------------------------------------------------
using bf16 = __nv_bfloat16;

template <typename T> __device__ __forceinline__ float to_f(T v);
template <> __device__ __forceinline__ float to_f<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f<bf16>(bf16 v) { return __bfloat162float(v); }

template <typename T> __device__ __forceinline__ T from_f(float v);
template <> __device__ __forceinline__ float from_f<float>(float v) { return v; }
template <> __device__ __forceinline__ bf16 from_f<bf16>(float v) { return __float2bfloat16_rn(v); }

// 8 consecutive elements <-> 8 floats (16 B for bf16, 32 B for fp32); p must be aligned.
template <typename T> __device__ __forceinline__ void load8(const T* p, float (&v)[8]);
template <> __device__ __forceinline__ void load8<float>(const float* p, float (&v)[8]) {
    float4 a = *reinterpret_cast<const float4*>(p);
    float4 b = *reinterpret_cast<const float4*>(p + 4);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
template <> __device__ __forceinline__ void load8<bf16>(const bf16* p, float (&v)[8]) {
    uint4 r = *reinterpret_cast<const uint4*>(p);
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&r);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        float2 f = __bfloat1622float2(h[i]);
        v[2 * i] = f.x; v[2 * i + 1] = f.y;
    }
}
template <typename T> __device__ __forceinline__ void store8(T* p, const float (&v)[8]);
template <> __device__ __forceinline__ void store8<float>(float* p, const float (&v)[8]) {
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
    *reinterpret_cast<float4*>(p + 4) = make_float4(v[4], v[5], v[6], v[7]);
}
template <> __device__ __forceinline__ void store8<bf16>(bf16* p, const float (&v)[8]) {
    uint4 r;
    __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&r);
#pragma unroll
    for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
    *reinterpret_cast<uint4*>(p) = r;
}

__device__ __forceinline__ float silu_f(float x) { return x / (1.0f + __expf(-x)); }
// exact-ish variant for the fp32 verification mode
__device__ __forceinline__ float silu_precise(float x) { return x / (1.0f + expf(-x)); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }

// cudaFuncAttributeMaxDynamicSharedMemorySize is a per-DEVICE attribute of a kernel: set it once per (kernel, device), so that
// a second GPU used from the same process gets it too
inline void ensure_dynamic_smem(const void* kern, int bytes) {
    static std::mutex mu;
    static std::set<std::pair<const void*, int>> done;
    int dev = 0;
    SYNT_CUDA(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lk(mu);
    if (done.count({kern, dev})) return;
    SYNT_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
    done.insert({kern, dev});
}

// ---- programmatic dependent launch (PDL) ----------------------------------------------
// The ~140 kernels of one sampling step form a chain.  Launched with the programmatic-stream-serialization attribute,
// kernel k+1 may be scheduled as soon as every CTA of kernel k has executed `griddepcontrol.launch_dependents` (first
// statement of every kernel here): its CTAs take the SMs that kernel k's tail frees, run their prologue (barrier
// init, TMEM allocation, descriptor prefetch) and block in `griddepcontrol.wait` until kernel k has completed and its
// writes are visible.  Rule for every kernel launched through launch_pdl: no global-memory access before pdl_wait(),
// and every CTA executes pdl_wait() (so "k+1 complete" implies "k complete" along the chain).
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_enter() { pdl_launch_dependents(); pdl_wait(); }

// SYNT_PDL: 0 = plain stream order, 1 = every kernel of the chain may launch early, 2 = only the GEMM kernels (their prologue --
// barrier init, TMEM allocation, tensor-map prefetch -- then overlaps the small GroupNorm-finalize kernel in front of them)
int pdl_mode();
inline bool pdl_enabled(bool gemm_kernel = false) { const int m = pdl_mode(); return m == 1 || (m == 2 && gemm_kernel); }

template <bool GEMM = false, typename... KArgs, typename... Args>
inline void launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, Args&&... args) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = pdl_enabled(GEMM) ? 1 : 0;
    cfg.attrs = at; cfg.numAttrs = 1;
    SYNT_CUDA(cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...));
}

}  // namespace synt
