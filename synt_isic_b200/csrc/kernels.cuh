// Launcher declarations for every kernel of the SYNT_ISIC hot path.
// Activations are NHWC ("pixels x channels"), storage type T = float (fp32 verification
// mode) or bf16 (production mode).  All launchers are asynchronous on `stream`.
#pragma once
#include "common.cuh"

namespace synt {

enum DType : int { DT_F32 = 0, DT_BF16 = 1 };
inline size_t dtype_size(int dt) { return dt == DT_F32 ? 4 : 2; }

// ---------------------------------------------------------------- convolution -------
// Implicit-GEMM convolution  out[m, n] = sum_k A[m, k] * Wt[n, k]  with
//   m = output pixel (b, oy, ox), n = output channel,
//   k = (tap, cin) over the KHxKW window of `in`, followed by optional 1x1 "shortcut"
//       segments read at the output pixel (times sc_stride) from sc0 / sc1.
// Epilogue: + bias[n] + bias2[n] + residual[m, n], optional ReLU.
struct ConvArgs {
    const void* in = nullptr;          // NHWC [B, H, W, Cin]
    int B = 0, H = 0, W = 0, Cin = 0;
    int KH = 3, KW = 3, stride = 1, pad = 1;
    int Ho = 0, Wo = 0, Cout = 0;
    // conv_tc2 only: second main source (channel concat, K order [tap][Cin + Cin1]) and fused input GroupNorm
    const void* in1 = nullptr; int Cin1 = 0;
    const float2* gn_ss = nullptr;     // [B][Cin + Cin1] (scale, shift)
    int gn_mode = 0;                   // 0 raw input, 1 GroupNorm affine, 2 affine + SiLU
    // conv_tc2 only: nearest-2x upsample fused into the 3x3 conv as four sub-pixel 2x2 convolutions on the
    // low-res input (2.25x fewer FLOPs): `in` is [B,H,W,Cin], `out` is [B,2H,2W,Cout], weight is the
    // phase-stacked matrix [4*Cout][4*Cin] (taps pre-summed); Ho/Wo describe the LOW-res tile grid here
    int up2x = 0;
    const void* sc0 = nullptr; int sc0_C = 0;    // NHWC [B, Ho*sc_stride, Wo*sc_stride, sc0_C]
    const void* sc1 = nullptr; int sc1_C = 0;
    int sc_stride = 1;
    const void* weight = nullptr;      // [Cout][Ktot] K-major; fp32 (SIMT) or bf16 (tcgen05)
    const float* bias = nullptr;       // [Cout]
    const float* bias2 = nullptr;      // [Cout] or null (time-embedding row)
    const void* residual = nullptr;    // NHWC [B, Ho, Wo, Cout] or null
    int relu = 0;
    void* out = nullptr;               // NHWC [B, Ho, Wo, Cout]
    float2* stats_out = nullptr;       // conv_tc2 only: per-channel (sum, sumsq) partials [B][slots][Cout] of `out`
    int ktot() const { return KH * KW * (Cin + Cin1) + sc0_C + sc1_C; }
};
// fp32-FMA implicit GEMM (verification mode and odd shapes); weights are fp32.
void conv_simt(const ConvArgs& a, int act_dtype, cudaStream_t s);
// tcgen05/TMEM/TMA implicit GEMM; activations and weights bf16, fp32 accumulate.
// Requires Cin, sc0_C, sc1_C multiples of 64 and Cout a multiple of 16.
bool conv_tc_supported(const ConvArgs& a);
void conv_tc(const ConvArgs& a, cudaStream_t s);
// persistent halo-tile kernel for 3x3 stride-1 convs (+ fused 1x1 shortcut segments), see conv_tc2.cu
bool conv_tc2_supported(const ConvArgs& a);
bool conv_tc2_is_k1(const ConvArgs& a);   // 1x1 projection on the no-halo four-slot variant (input transform once per load)
void conv_tc2(const ConvArgs& a, cudaStream_t s);
int conv_tc2_stats_slots(const ConvArgs& a);     // partial rows per image written when stats_out != null
void conv_tc2_set_pair(int on);                  // experimental CTA-pair (cta_group::2) variant of the N = 128 kernel
// experimental: 3x3 stride-1 conv with one halo-tile load per channel chunk (see conv_tc_halo.cu)
void conv_tc_halo(const ConvArgs& a, int variant, cudaStream_t s);

// UNet conv_in: x fp32 NCHW [B,3,H,W] -> NHWC T [B,H,W,64], 3x3 pad 1.
struct ConvInW { float w[27][64]; float b[64]; };          // k = tap*3 + c
// stats_out (optional, only when conv_in3_stats_slots(H, W, dt) > 0): per-channel (sum, sumsq) partial rows
// [B][slots][64] of the stored output, for the first GroupNorm
int conv_in3_stats_slots(int H, int W, int dt);
void conv_in3(const float* x_nchw, const ConvInW& w, int B, int H, int W, void* out, int dt, float2* stats_out,
              cudaStream_t s);

// tcgen05 version (conv_in_tc.cu, bf16 mode, 128x128 images): w_taps from conv_in_tc_pack_weights (device copy), same
// statistics layout as conv_in3 ([B][16][64] partial rows, one per 8-row band)
void conv_in_tc(const float* x_nchw, int B, const void* w_taps, const float* bias, void* out, float2* stats, cudaStream_t s);
void conv_in_tc_pack_weights(const ConvInW& w, uint16_t* out);
int conv_in_tc_weight_bytes();

// UNet tail: eps = conv_out(SiLU(GN(h))) fused with DDPMScheduler.step.
struct ConvOutW { float w[9][64][3]; float b[3]; };
struct SchedArgs {
    // when x != null the scheduler update runs in the epilogue:
    //   x0 = clamp((x - sqrt_b*eps)/sqrt_a, -1, 1); x' = c_x0*x0 + c_xt*x + sigma*z
    float* x = nullptr;                // fp32 NCHW [B,3,H,W], updated IN PLACE
    const float* coef = nullptr;       // device [5] = {sqrt_b, sqrt_a, c_x0, c_xt, sigma}
    const float* z = nullptr;          // injected noise of step 0 for these images, or null -> Philox
    long long z_step_stride = 0;       // elements between consecutive steps of z
    unsigned long long seed = 0;       // Philox key (used when z == null)
    const int* step_ptr = nullptr;     // device step counter (selects z / traj / eps frame, Philox offset)
    float* traj = nullptr;             // optional copy of x' (trajectory frame of step 0)
    long long traj_step_stride = 0;
    long long eps_step_stride = 0;     // stride of the eps tap between steps
    long long image_offset = 0;        // global image index of image 0 of this call (Philox stream id)
    const long long* image_keys = nullptr;   // optional device [B]: Philox stream id of image b (replaces image_offset + b),
                                       // so that an image's noise depends on ITS key only, not on the batch it is sampled in
    // coalition decoding (permutation Time-SHAP over denoising steps, README.md:171-221 of the reference): image b takes
    // the transition of step s only if step_mask[s * mask_stride + b] != 0, otherwise x_b is frozen for that step;
    // noise_shared: every image of the batch draws the SAME Philox noise field (common random numbers across coalitions)
    const unsigned char* step_mask = nullptr;
    int mask_stride = 0;
    int noise_shared = 0;
};
// bfrag (bf16 mode, nullable): conv_out weights as per-lane mma.sync m16n8k16 B fragments, uint2[36 k-steps][32 lanes]
// (k-step = tap*4 + 16-channel block; lane (g = lane/4, t = lane%4): .x = w[k0 + 2t, 2t+1][n = g], .y = same at k0 + 8)
void conv_out3(const void* h, int dt, const float2* scale_shift, const ConvOutW& w, const void* bfrag, int B, int H, int W,
               float* eps_nchw /*nullable*/, const SchedArgs& sch, cudaStream_t s);

// Stand-alone DDPMScheduler.step on fp32 NCHW tensors (the drop-in scheduler object).
void ddpm_step(const float* eps, const float* x, const float* z, float* out, long long n, float sqrt_b,
               float sqrt_a, float c_x0, float c_xt, float sigma, cudaStream_t s);

// ---------------------------------------------------------------- GroupNorm ---------
// x = concat_channels(src0[C0], src1[C1]); statistics per (b, group) over HW x (C/G).
//   gn_stats    -> partial (sum, sumsq) per (b, chunk, group)
//   gn_finalize -> per (b, c): scale = rstd*gamma, shift = beta - mean*rstd*gamma
//   gn_apply    -> out = act(x*scale + shift) (NHWC T, channels concatenated)
int gn_num_chunks(int B, int HW);
void gn_stats(const void* src0, int C0, const void* src1, int C1, int dt, int B, int HW, int G, float2* partials,
              int nchunk, cudaStream_t s);
void gn_finalize(const float2* partials, int B, int nchunk, int G, int C, int HW, float eps, const float* gamma,
                 const float* beta, float2* scale_shift, cudaStream_t s);
// finalize from PER-CHANNEL partials of one or two (concatenated) tensors: partA [B][SA][C0], partB [B][SB][C1]
void gn_finalize_channels(const float2* partA, int SA, int C0, const float2* partB, int SB, int C1, int B, int G, int HW,
                          float eps, const float* gamma, const float* beta, float2* scale_shift, cudaStream_t s);
// out = act(GroupNorm(concat(src0, src1))) with the statistics finalize folded into the kernel:
// part0/part1 are the per-channel (sum, sumsq) partial rows [B][slots][C] of the two tensors
void gn_apply_fused(const void* src0, const float2* part0, int slots0, int C0, const void* src1, const float2* part1,
                    int slots1, int C1, int dt, int B, int HW, int G, float eps, const float* gamma, const float* beta,
                    const float2* scale_shift /* precomputed by gn_finalize_channels, or null = finalize in-kernel */,
                    int silu, void* out, cudaStream_t s);

// ---------------------------------------------------------------- misc --------------
void upsample_nearest2x(const void* in, int dt, int B, int H, int W, int C, void* out, cudaStream_t s);
// softmax(q k^T / sqrt(8)) v over qkv NHWC-token layout [B, N, 3*C] (q | k | v), heads of 8.
void attention_simt(const void* qkv, int dt, int B, int N, int C, void* out, cudaStream_t s);
bool attention_tc_supported(int N, int C);
// tcgen05 attention: qkv [B, N, 3C] bf16 = (q | k | v) with q pre-scaled by log2(e)/sqrt(8);
// vt_scratch is unused (the padded V^T operand is built in shared memory by the kernel) and may be null.
void attention_tc(const void* qkv, int B, int N, int C, void* vt_scratch, void* out, cudaStream_t s);

// time embedding tables: emb[t] = Linear2(SiLU(Linear1(sincos(t)))) for t in [0,T),
// temb[t][off_r + c] = Linear_r(SiLU(emb[t]))[c] for every resnet r (sum Cout = ntot).
void time_embed_table(const float* freqs32, const float* w1, const float* b1, const float* w2, const float* b2,
                      int T, float* emb_silu /*[T][256]*/, cudaStream_t s);
void time_proj_table(const float* emb_silu, const float* w /*[ntot][256]*/, const float* b, int T, int ntot,
                     float* table /*[T][ntot]*/, cudaStream_t s);
// copy row `t` (or timesteps[*step_ptr]) of the tables into the fixed per-step buffers
void select_timestep(const float* table, int ntot, const float* coef_table /*[T][5] or null*/,
                     const int* timesteps, const int* step_ptr, int t_direct, float* temb_cur, float* coef_cur,
                     cudaStream_t s);
void advance_step(int* step_ptr, cudaStream_t s);
void set_step(int* step_ptr, int value, cudaStream_t s);   // *step_ptr = value (kernel argument: no host buffer to keep alive)

// debug / format helpers
void nhwc_to_nchw_f32(const void* in, int dt, int B, int HW, int C, float* out, cudaStream_t s);
void to_uint8_hwc(const float* x_nchw, int B, int H, int W, int mode, unsigned char* out, cudaStream_t s);

// ---------------------------------------------------------------- classifier --------
// clamp((x+1)/2) -> bilinear 128->224 (align_corners=False) -> ImageNet normalise; optional
// fused intervention blend x~ = clamp(x(1-M) + I M, -1, 1) before it.
void classifier_preprocess(const float* x_nchw, int B, int Hin, int Win, int Hout, int Wout, void* out_nhwc,
                           int Cpad, int dt, cudaStream_t s);
void maxpool3x3s2(const void* in, int dt, int B, int H, int W, int C, void* out, cudaStream_t s);
void avgpool_fc(const void* in, int dt, int B, int HW, int C, const float* w, const float* b, int nout,
                float* logits, cudaStream_t s);
// ---- input-gradient path (resnet_grad.cu): d log(p_c + 1e-8) / dx for Integrated Gradients, xai/XAI.py:1039-1109 ----
void score_grad(const float* logits, int B, int nc, int target, float* score /*nullable*/, float* dlogits, cudaStream_t s);
void avgpool_fc_bwd(const float* dlogits, const float* w, const void* feat, int dt, int B, int HW, int C, int nc, void* out,
                    cudaStream_t s);                                           // ReLU mask (feat > 0) fused
void relu_mask(const void* g, const void* act, int dt, long long n, void* out /* may alias g */, cudaStream_t s);
void zero_insert2x(const void* g, const void* act /*nullable ReLU mask*/, int dt, int B, int Ho, int Wo, int C, void* out,
                   cudaStream_t s);
// max-pool that also records the winning window element (code = dy*3 + dx, one byte per output element) and its adjoint
void maxpool3x3s2_idx(const void* in, int dt, int B, int H, int W, int C, void* out, unsigned char* idx, cudaStream_t s);
void maxpool3x3s2_bwd(const void* dpool, const unsigned char* idx, const void* act, int dt, int B, int H, int W, int C, void* dact,
                      cudaStream_t s);
void stem_dgrad(const void* g, int dt, int B, const float* w /*[64][147] fp32*/, float* dpre /*[B,224,224,3]*/, cudaStream_t s);
void classifier_preprocess_bwd(const float* dpre, const float* x, int B, int Hin, int Win, int Hout, int Wout, float* dx,
                               cudaStream_t s);
void dgrad_weights(const void* src, int src_ld, int src_off, int cin_f, int cout_f, int taps, int bf, void* dst, int dst_ld,
                   int dst_off, cudaStream_t s);
void ig_interpolate(const float* x, const float* base, int n_steps, long long per, float* out, cudaStream_t s);
void ig_reduce(const float* grads, const float* x, const float* base, int n_steps, long long per, float* out, cudaStream_t s);
// interventions (xai/XAI.py:1495-1575): type 0=zero 1=mean 2=blur5 3=noise(injected) 4=given tensor
// blur_k: odd box size of type 2 (the reference's blur_kernel kwarg, default 5); interv_out (nullable) receives the
// intervention tensor I itself (the reference returns it and reports mean |I| as intervention_strength, XAI.py:1584,1592)
void intervene_blend(const float* x, const float* mask, const float* aux, int type, float noise_std, int blur_k, int B, int C,
                     int H, int W, float* out, float* interv_out, cudaStream_t s);
void patch_mask_apply(const float* x, const unsigned char* patch_masks, int n_masks, int C, int H, int W,
                      int patch, float* out, cudaStream_t s);

}  // namespace synt
