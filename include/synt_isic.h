/* synt_isic.h -- C ABI of libsynt_isic_b200.so (sm_100a only, no CPU fallback).
 *
 * The reference (fims9000/SYNT_ISIC) has NO native interface: its hot path is two lines of
 * Python that call third-party PyTorch modules,
 *
 *     noise_pred = model(latents, t).sample                         image_generator.py:400
 *     latents = scheduler.step(noise_pred, t, latents).prev_sample  image_generator.py:403
 *     logits  = self.model(self.preprocess_for_classifier(x))       xai/XAI.py:433-436
 *
 * so the drop-in boundary is the Python object protocol, and this
 * header is the C boundary underneath it: plain pointers and sizes, no torch types.  Each
 * entry point cites the reference call it replaces.  INTEGRATION.md shows the ctypes stub.
 *
 * Conventions: every function returns 0 on success or a negative code (-1 bad argument,
 * -2 CUDA error, -3 unsupported, -4 driver entry point missing); synt_last_error() returns
 * the message of the calling thread's last failure.  Functions never throw.  `stream` is a
 * cudaStream_t passed as void* (NULL = legacy default stream).  Device pointers must belong
 * to the current CUDA device.  Handles are thread-compatible, not thread-safe (the reference
 * drives them from one worker thread, main.py:43-56).  Images are fp32 NCHW [B,3,128,128] in
 * [-1,1] exactly as the reference's tensors (image_generator.py:379).
 */
#ifndef SYNT_ISIC_H
#define SYNT_ISIC_H
#ifdef __cplusplus
extern "C" {
#endif

#define SYNT_DTYPE_F32 0   /* fp32 verification mode: fp32 FMA kernels end to end          */
#define SYNT_DTYPE_BF16 1  /* production mode: bf16 tcgen05 GEMMs, fp32 accumulate/state   */

typedef struct synt_unet synt_unet_t;
typedef struct synt_resnet18 synt_resnet18_t;

const char* synt_version(void);
const char* synt_last_error(void);

/* ---------------- UNet2DModel  (core/generator/model_manager.py:173-194) ---------------- */
/* Parameter manifest in diffusers state_dict naming (load_state_dict(strict=True)
 * at xai/XAI.py:605).  The host packs the fp32 tensors back to back in manifest order. */
int synt_unet_num_params(void);
int synt_unet_param_info(int index, char* name, int name_cap, long long* numel, long long* offset);
long long synt_unet_total_param_count(void);            /* == 25,304,963 */

/* replaces ModelManager._create_model_architecture + load_state_dict (model_manager.py:138-143) */
int synt_unet_create(const float* params_host, long long n_params, int dtype, synt_unet_t** out);
int synt_unet_destroy(synt_unet_t* h);

/* replaces `model(latents, t).sample`  (image_generator.py:400): eps = UNet(x, t) */
int synt_unet_forward(synt_unet_t* h, const float* x_dev, int B, int t, float* eps_dev, void* stream);
/* same forward, additionally copies the output of module `tap` (diffusers module path, e.g.
 * "down_blocks.0.resnets.1") as fp32 NCHW into out_dev; used by the parity tests only */
int synt_unet_debug_forward(synt_unet_t* h, const float* x_dev, int B, int t, const char* tap, float* out_dev,
                            long long out_cap, int* C, int* H, int* W, void* stream);

/* replaces scheduler.set_timesteps(n) + the per-step coefficient algebra of DDPMScheduler.step
 * (model_manager.py:199-209): coef[i] = {sqrt(1-abar_t), sqrt(abar_t), c_x0, c_xt, sigma} */
int synt_unet_set_schedule(synt_unet_t* h, int n_steps, const int* timesteps_host, const float* coef_host);
/* replaces the whole loop image_generator.py:395-403 for steps [step_begin, step_end):
 *   x is updated in place; z_dev (optional) is injected noise [n_steps][B,3,H,W], otherwise
 *   Philox(seed, image_offset+b, step); traj_dev (optional) receives x after every step
 *   (image_generator.py:407), eps_tap_dev (optional) the raw eps of every step.
 *   micro_batch <= 0 selects the library default; use_graph != 0 replays a captured CUDA graph. */
int synt_unet_sample(synt_unet_t* h, float* x_dev, int B, const float* z_dev, unsigned long long seed,
                     long long image_offset, float* traj_dev, float* eps_tap_dev, int step_begin, int step_end,
                     int micro_batch, int use_graph, void* stream);
/* Coalition decoding for the permutation Time-SHAP over denoising steps (README.md:171-221 of the reference; no code
 * in the reference): with mask_dev != NULL (uint8 [n_steps][B], device) image b takes the transition of step s of the
 * following synt_unet_sample calls only if mask[s*B + b] != 0 and is frozen otherwise; noise_shared != 0 makes every
 * image of the batch draw the same in-kernel noise field (common random numbers across coalitions).  NULL / 0 resets. */
int synt_unet_set_step_mask(synt_unet_t* h, const unsigned char* mask_dev, int noise_shared);
/* Per-image noise streams: with keys_dev != NULL (int64 [B], device; must stay valid while sampling) image b of the following
 * synt_unet_sample calls draws its in-kernel noise from Philox(seed, stream = keys_dev[b], step) instead of stream =
 * image_offset + b, so the image produced for a key does not depend on the batch it is sampled in (the reference draws
 * fresh noise per image and step, image_generator.py:403).  NULL resets. */
int synt_unet_set_image_keys(synt_unet_t* h, const long long* keys_dev);
/* host-buffer form of the same call (what a non-PyTorch caller binds): H2D of x_T, all steps of
 * the current schedule, uint8 HWC conversion (image_generator.py:441-447), D2H of the images.
 * x_final_host (optional) receives the fp32 result. */
int synt_unet_generate_host(synt_unet_t* h, const float* xT_host, int B, unsigned long long seed,
                            long long image_offset, int micro_batch, unsigned char* images_u8_host,
                            float* x_final_host);
/* measurement hook: runs ONE eager sampling step with a CUDA-event pair around every kernel launch and
 * returns, per category (0 conv_tcgen05, 1 conv_fp32, 2 groupnorm_stats, 3 groupnorm_apply, 4 attention,
 * 5 upsample, 6 conv_in, 7 conv_out+scheduler, 8 misc; arrays of 16), the summed device time [ms], the
 * algorithmic FLOPs and the launch count.  x is advanced by one step. */
int synt_unet_profile_step(synt_unet_t* h, float* x_dev, int B, int micro_batch, double* ms_out, double* flops_out,
                           int* launches_out, void* stream);
/* per-launch rows {category, ms, flops, M, N, K} of the last profiled step; returns the row count */
int synt_unet_profile_records(synt_unet_t* h, double* rows6, int cap);
long long synt_unet_workspace_bytes(synt_unet_t* h);
long long synt_unet_launch_count(synt_unet_t* h);        /* kernels launched since creation */

/* ---------------- DDPMScheduler  (core/generator/model_manager.py:196-226) -------------- */
/* schedule 0 = "squaredcos_cap_v2", 1 = "linear"(beta_start, beta_end); host-only arithmetic.
 * timesteps_out[n_steps] (int64, "leading" spacing), coef_out[n_steps][5], alphas_cumprod_out[num_train] */
int synt_ddpm_tables(int num_train, int schedule, float beta_start, float beta_end, int n_steps,
                     long long* timesteps_out, float* coef_out, float* alphas_cumprod_out);
/* replaces `scheduler.step(noise_pred, t, latents).prev_sample` (image_generator.py:403) */
int synt_ddpm_step(const float* eps_dev, const float* x_dev, const float* z_dev, float* out_dev, long long n,
                   const float* coef5_host, void* stream);
/* replaces image_generator.py:441-447 (mode 0) / diffusion_generator.py:147-148 (mode 1) */
int synt_to_uint8(const float* x_dev, int B, int H, int W, int mode, unsigned char* out_dev, void* stream);

/* ---------------- MelanomaClassifierAdaptive / ResNet18  (xai/XAI.py:357-471) ----------- */
int synt_resnet18_num_params(void);
int synt_resnet18_param_info(int num_classes, int index, char* name, int name_cap, long long* numel,
                             long long* offset);   /* torchvision resnet18 state_dict names */
int synt_resnet18_create(const float* params_host, long long n_params, int num_classes, int dtype,
                         synt_resnet18_t** out);
int synt_resnet18_destroy(synt_resnet18_t* h);
/* replaces `self.model(self.preprocess_for_classifier(x))` (XAI.py:399-436): x in [-1,1] */
int synt_resnet18_logits(synt_resnet18_t* h, const float* x_dev, int B, float* logits_dev, void* stream);
int synt_resnet18_logits_host(synt_resnet18_t* h, const float* x_host, int B, float* logits_host);
int synt_resnet18_debug(synt_resnet18_t* h, const float* x_dev, int B, const char* tap, float* out_dev,
                        long long out_cap, int* C, int* H, int* W, void* stream);
long long synt_resnet18_launch_count(synt_resnet18_t* h);
/* measurement hook: one warm forward of B <= 512 images with CUDA events between the front end (preprocess + 7x7 stem +
 * max-pool), the 16 body convolutions and the pool + FC head; ms_out[3] */
int synt_resnet18_profile(synt_resnet18_t* h, const float* x_dev, int B, float* logits_dev, double* ms_out, void* stream);

/* Input gradient of the per-class score s = log(softmax(logits)[target_class] + 1e-8) (get_per_class_score,
 * XAI.py:443-459): replaces the autograd pass that captum's IntegratedGradients (XAI.py:1039-1085) and the plain gradient
 * attribution (XAI.py:1087-1109) run through the classifier.  score_dev [B] (nullable), grad_dev [B,3,128,128] fp32. */
int synt_resnet18_score_grad(synt_resnet18_t* h, const float* x_dev, int B, int target_class, float* score_dev,
                             float* grad_dev, void* stream);
/* test hook: gradient at an intermediate tensor ("grad:layer4.1" ... "grad:layer1.0" = at the block's pre-ReLU output,
 * "grad:maxpool", "grad:relu" (stem output, ReLU-masked), "grad:preprocess"), NCHW fp32; B <= 32 */
int synt_resnet18_grad_debug(synt_resnet18_t* h, const float* x_dev, int B, int target_class, const char* tap,
                             float* out_dev, long long out_cap, int* C, int* H, int* W, void* stream);
/* Integrated-Gradients path points and reduction (captum method='riemann_right', XAI.py:1069-1074):
 *   out[k] = baseline + (k+1)/n_steps * (x - baseline), k = 0..n_steps-1            out [n_steps][per_image]
 *   out    = (x - baseline) * sum_k grads[k] / n_steps                               out [per_image]            */
int synt_ig_interpolate(const float* x_dev, const float* baseline_dev, int n_steps, long long per_image, float* out_dev,
                        void* stream);
int synt_ig_reduce(const float* grads_dev, const float* x_dev, const float* baseline_dev, int n_steps, long long per_image,
                   float* out_dev, void* stream);

/* replaces counterfactual_intervention_advanced (XAI.py:1454-1597):
 *   out = clamp(x*(1-M) + I*M, -1, 1); type 0 zero, 1 per-channel mean, 2 5x5 box blur,
 *   3 noise (aux = injected N(0,1) tensor, scaled by noise_std), 4 aux is the intervention itself */
int synt_intervene_blend(const float* x_dev, const float* mask_dev, const float* aux_dev, int type, float noise_std,
                         int B, int C, int H, int W, float* out_dev, void* stream);
/* the same with the reference's `blur_kernel` kwarg (odd box size of type 2, XAI.py:1474,1511-1527; the plain call uses 5)
 * and, when intervention_out_dev != NULL, the intervention tensor I itself [B,C,H,W] (result key 'intervention' and
 * statistics['intervention_strength'] = mean |I|, XAI.py:1584,1592) */
int synt_intervene_blend_ex(const float* x_dev, const float* mask_dev, const float* aux_dev, int type, float noise_std,
                            int blur_kernel, int B, int C, int H, int W, float* out_dev, float* intervention_out_dev,
                            void* stream);
/* replaces the masking loop of compute_shap_approximation (XAI.py:1143-1161):
 *   out[i] = x with every patch whose mask byte is 0 set to 0; patch_masks [n][H/p][W/p] */
int synt_patch_mask_apply(const float* x_dev, const unsigned char* patch_masks_dev, int n_masks, int C, int H, int W,
                          int patch, float* out_dev, void* stream);

/* replaces select_regions_advanced (XAI.py:1340-1451; numpy percentile + scipy.ndimage closing x2 / opening / label /
 * small-component removal) for n_maps attribution maps at once, one CTA per map, H*W <= 16384:
 *   attr_dev [n_maps][C][H][W] fp32 (saliency = channel L2 norm) or, with use_abs, [n_maps][H][W] (saliency = |x|);
 *   top-k (bottom = 0: saliency >= percentile(100 - k)) or bottom-k (bottom = 1: saliency <= percentile(k));
 *   mask_dev [n_maps][H][W] bytes 0/1; stats_dev [n_maps][8] fp64 = selected pixels, threshold, mean, std of the saliency,
 *   mean, std, max, min of the saliency over the selection (0 when nothing is selected) */
int synt_select_regions(const float* attr_dev, int n_maps, int C, int H, int W, int use_abs, double k_percent, int bottom,
                        int morphology_cleanup, int connectivity, unsigned char* mask_dev, double* stats_dev, void* stream);

/* replaces the resampling loops of statistical_validation_comprehensive (XAI.py:1708-2005), one thread per replicate, float64
 * means in numpy's pairwise summation order:
 *   bootstrap   (XAI.py:1852-1878): out[b] = mean(top[idx_top[b][0..n1)]) - mean(bottom[idx_bottom[b][0..n2)])
 *   permutation (XAI.py:1882-1913): out[b] = mean(combined[perm[b][0..n1)]) - mean(combined[perm[b][n1..n)])
 * idx / perms (int32, device) inject the reference's numpy draws (bit-identical replicates); NULL = in-kernel Philox(seed)
 * draws (bootstrap: uniform indices; permutation: Fisher-Yates, n <= 1024). */
int synt_stat_bootstrap_mean_diff(const double* top_dev, int n1, const double* bottom_dev, int n2, const int* idx_top_dev,
                                  const int* idx_bottom_dev, unsigned long long seed, int n_bootstrap, double* out_dev,
                                  void* stream);
int synt_stat_permutation_mean_diff(const double* combined_dev, int n, int n1, const int* perms_dev, unsigned long long seed,
                                    int n_permutations, double* out_dev, void* stream);

/* ---------------- test hook (kernel-level parity tests; not a reference entry point) ------ */
/* one convolution on caller-provided NHWC tensors: use_tc=1 tcgen05 (bf16), 0 fp32-FMA carrier.
 * weight is K-major [Cout][K*K*Cin + sc0_C + sc1_C], bf16 for use_tc else fp32. */
int synt_debug_conv(int use_tc, int act_dtype, const void* in_dev, int B, int H, int W, int Cin, int K, int stride,
                    int pad, const void* sc0_dev, int sc0_C, const void* sc1_dev, int sc1_C, int sc_stride,
                    const void* weight_dev, const float* bias_dev, const float* bias2_dev, const void* residual_dev,
                    int relu, void* out_dev, int Cout, void* stream);

/* the persistent conv kernel with its fused-input features: main input = channel concat (in | in1),
 * GroupNorm scale/shift gn_ss [B][Cin+Cin1] (float2) applied in-kernel (gn_mode 1 affine, 2 affine+SiLU);
 * stats_out (optional) receives per-channel (sum, sumsq) partial rows [B][*stats_slots][Cout] of the output */
int synt_debug_conv_gn(const void* in_dev, int Cin, const void* in1_dev, int Cin1, const void* gn_ss_dev, int gn_mode,
                       int B, int H, int W, int K, const void* sc0_dev, int sc0_C, const void* weight_dev,
                       const float* bias_dev, const void* residual_dev, void* out_dev, int Cout, void* stats_out_dev,
                       int* stats_slots, void* stream);
/* Upsample2D(nearest 2x) + conv3x3 through the fused sub-pixel path (four 2x2 convolutions on the low-res
 * input): in [B,H,W,Cin] bf16 (device), w_host the raw [Cout][Cin][3][3] fp32 filter (HOST), out [B,2H,2W,Cout]. */
int synt_debug_conv_up2x(const void* in_dev, int B, int H, int W, int Cin, const float* w_host, const float* bias_dev,
                         void* out_dev, int Cout, void* stats_out_dev, int* stats_slots, void* stream);
/* Experimental: run the N = 128 conv_tc2 launches as CTA pairs (tcgen05 cta_group::2, M = 256 per MMA); off by default. */
int synt_debug_set_conv_pair(int on);
/* Host-only (no GPU needed): phase-stacked filter of the fused Upsample2D + conv3x3: w [Cout][Cin][3][3] ->
 * out [4*Cout][4*Cin], row = (py*2+px)*Cout + n, col = (ty*2+tx)*Cin + c; output pixel (2y+py, 2x+px) =
 * sum over the 2x2 low-res window rows y+py-1+ty, columns x+px-1+tx. */
int synt_debug_pack_upsample_phases(const float* w_host, int Cout, int Cin, float* out_host);
/* softmax(q k^T / sqrt(8)) v on caller-provided tensors qkv [B,N,3C] (q|k|v), out [B,N,C]; use_tc=1: the tcgen05
 * kernel, bf16, q already multiplied by log2(e)/sqrt(8). */
int synt_debug_attention(int use_tc, int act_dtype, const void* qkv_dev, int B, int N, int C, void* out_dev,
                         void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SYNT_ISIC_H */
