// Classifier preprocess of ONE output pixel (xai/XAI.py:399-431): clamp((x+1)/2, 0, 1) -> bilinear 128 -> 224
// (align_corners=False; == antialias=True when upsampling) -> ImageNet normalise.  Shared by the stand-alone preprocess kernel
// and the fused front ends (resnet.cu, stem_tc.cu).
#pragma once
#include "common.cuh"

namespace synt {

// =============================================================== kernels ============
// classifier preprocess of one output pixel (xai/XAI.py:399-431): clamp((x+1)/2, 0, 1) -> bilinear Hin x Win -> Hout x Wout
// (align_corners=False; antialias is a no-op when upsampling) -> ImageNet normalise.  One definition, used by the
// stand-alone kernel and by the fused stem, so both produce bit-identical values.
__device__ __forceinline__ void preprocess_pixel(const float* __restrict__ img /* [3][Hin][Win] */, int Hin, int Win, float sy, float sx,
                                                 int oy, int ox, float (&v3)[3]) {
    const float mean[3] = {0.485f, 0.456f, 0.406f}, stdv[3] = {0.229f, 0.224f, 0.225f};
    float fy = ((float)oy + 0.5f) * sy - 0.5f; if (fy < 0.f) fy = 0.f;
    float fx = ((float)ox + 0.5f) * sx - 0.5f; if (fx < 0.f) fx = 0.f;
    const int y0 = (int)fy, x0 = (int)fx;
    const int y1 = y0 + (y0 < Hin - 1 ? 1 : 0), x1 = x0 + (x0 < Win - 1 ? 1 : 0);
    const float ly = fy - (float)y0, lx = fx - (float)x0;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const float* pl = img + (long long)c * Hin * Win;
        auto px = [&](int yy, int xx) {
            float v = (pl[yy * Win + xx] + 1.0f) / 2.0f;
            return fminf(fmaxf(v, 0.f), 1.f);
        };
        const float top = px(y0, x0) * (1.f - lx) + px(y0, x1) * lx;
        const float bot = px(y1, x0) * (1.f - lx) + px(y1, x1) * lx;
        const float v = top * (1.f - ly) + bot * ly;
        v3[c] = (v - mean[c]) / stdv[c];
    }
}

}  // namespace synt
