// Knock-out timing of the attention kernel (which role / which work bounds it?): compiles csrc/attention_tc.cu with
// -DATC_KNOCK=<mask> (or any other -D knob of that file) into a stand-alone binary and times the B=64 launch.
//   tools/ubench/build_att.sh k1 -DATC_KNOCK=1 k4 -DATC_KNOCK=4 ...   ->  tools/ubench/att_k1 [N]
#include "../../synt_isic_b200/csrc/attention_tc.cu"
#include <cstdlib>
namespace synt { int pdl_mode() { return 0; } }
int main(int argc, char** argv) {
    const int N = argc > 1 ? atoi(argv[1]) : 1024, B = argc > 2 ? atoi(argv[2]) : 64, C = 256;
    const size_t n = (size_t)B * N * 3 * C;
    std::vector<__nv_bfloat16> h(n);
    uint32_t st = 12345u;
    for (size_t i = 0; i < n; ++i) {           // roughly N(0, 0.7)
        float a = 0.f;
        for (int k = 0; k < 4; ++k) { st = st * 1664525u + 1013904223u; a += (st >> 8) * (1.0f / 16777216.0f) - 0.5f; }
        h[i] = __float2bfloat16(a * 1.2f);
    }
    __nv_bfloat16 *qkv, *out;
    cudaMalloc(&qkv, n * 2); cudaMalloc(&out, (size_t)B * N * C * 2);
    cudaMemcpy(qkv, h.data(), n * 2, cudaMemcpyHostToDevice);
    try {
        for (int i = 0; i < 3; ++i) synt::attention_tc(qkv, B, N, C, nullptr, out, 0);
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        cudaEventRecord(e0);
        for (int i = 0; i < 10; ++i) synt::attention_tc(qkv, B, N, C, nullptr, out, 0);
        cudaEventRecord(e1);
        cudaError_t e = cudaDeviceSynchronize();
        float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
        printf("ATC_KNOCK=%d N=%d: %.1f us per launch (%s)\n", ATC_KNOCK, N, ms * 100.0f, cudaGetErrorString(e));
    } catch (const std::exception& ex) { printf("error: %s\n", ex.what()); return 1; }
    return 0;
}
