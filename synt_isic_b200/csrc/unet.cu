// UNet2D execution plan + DDPM sampling loop behind the C ABI (include/synt_isic.h).
//
// What it replaces in the reference (paths under /root/reference):
//   UNet2DModel(...) construction           core/generator/model_manager.py:173-194
//   model(latents, t).sample                core/generator/image_generator.py:400
//   scheduler.step(...).prev_sample         core/generator/image_generator.py:403
//   the T-step loop                         core/generator/image_generator.py:395-403
// The network topology is diffusers' UNet2DModel for those constructor arguments
// (SURVEY.md Appendix A.1); weights arrive in diffusers state_dict naming (A.2).
#include "kernels.cuh"
#include "pool.h"
#include "../../include/synt_isic.h"
#include <cmath>
#include <cstring>
#include <functional>
#include <memory>
#include <string>
#include <vector>

namespace synt {

thread_local std::string g_last_error;

// ------------------------------------------------------------------ manifest ----------
struct ParamSpec { std::string name; long long numel; long long offset; };

static const int kBlockCh[4] = {64, 128, 256, 256};
static const bool kDownAttn[4] = {false, false, true, false};
static const bool kUpAttn[4] = {false, true, false, false};
constexpr int kTembDim = 256, kGroups = 32, kImg = 128, kTrainSteps = 1000;
constexpr float kGnEps = 1e-5f;

struct ResnetCfg { std::string prefix; int cin, cout; };
struct AttnCfg { std::string prefix; int C; };

struct Manifest {
    std::vector<ParamSpec> params;
    long long total = 0;
    void add(const std::string& n, long long numel) { params.push_back({n, numel, total}); total += numel; }
    void conv(const std::string& p, int cout, int cin, int k) { add(p + ".weight", (long long)cout * cin * k * k); add(p + ".bias", cout); }
    void linear(const std::string& p, int out, int in) { add(p + ".weight", (long long)out * in); add(p + ".bias", out); }
    void norm(const std::string& p, int c) { add(p + ".weight", c); add(p + ".bias", c); }
    void resnet(const std::string& p, int cin, int cout) {
        norm(p + ".norm1", cin); conv(p + ".conv1", cout, cin, 3); linear(p + ".time_emb_proj", cout, kTembDim);
        norm(p + ".norm2", cout); conv(p + ".conv2", cout, cout, 3);
        if (cin != cout) conv(p + ".conv_shortcut", cout, cin, 1);
    }
    void attn(const std::string& p, int c) {
        norm(p + ".group_norm", c); linear(p + ".to_q", c, c); linear(p + ".to_k", c, c); linear(p + ".to_v", c, c);
        linear(p + ".to_out.0", c, c);
    }
    long long find(const std::string& n) const {
        for (auto& s : params) if (s.name == n) return s.offset;
        throw Error(-1, "manifest: no parameter named " + n);
    }
};

// up-block resnet input channels (diffusers UpBlock2D): see SURVEY.md A.1
static void up_resnet_channels(int i, int j, int& c_h, int& c_skip, int& cout) {
    const int rev[4] = {256, 256, 128, 64};
    cout = rev[i];
    const int prev = rev[i == 0 ? 0 : i - 1];
    const int inp = rev[i + 1 < 4 ? i + 1 : 3];
    c_skip = (j == 2) ? inp : cout;
    c_h = (j == 0) ? prev : cout;
}

static Manifest build_manifest() {
    Manifest m;
    m.conv("conv_in", 64, 3, 3);
    m.linear("time_embedding.linear_1", kTembDim, 64);
    m.linear("time_embedding.linear_2", kTembDim, kTembDim);
    int out = 64;
    for (int i = 0; i < 4; ++i) {
        const int inp = out; out = kBlockCh[i];
        const std::string b = "down_blocks." + std::to_string(i);
        for (int j = 0; j < 2; ++j) m.resnet(b + ".resnets." + std::to_string(j), j == 0 ? inp : out, out);
        if (kDownAttn[i]) for (int j = 0; j < 2; ++j) m.attn(b + ".attentions." + std::to_string(j), out);
        if (i != 3) m.conv(b + ".downsamplers.0.conv", out, out, 3);
    }
    m.attn("mid_block.attentions.0", 256);
    m.resnet("mid_block.resnets.0", 256, 256);
    m.resnet("mid_block.resnets.1", 256, 256);
    for (int i = 0; i < 4; ++i) {
        const std::string b = "up_blocks." + std::to_string(i);
        int cout = 0;
        for (int j = 0; j < 3; ++j) {
            int ch, cs; up_resnet_channels(i, j, ch, cs, cout);
            m.resnet(b + ".resnets." + std::to_string(j), ch + cs, cout);
        }
        if (kUpAttn[i]) for (int j = 0; j < 3; ++j) m.attn(b + ".attentions." + std::to_string(j), cout);
        if (i != 3) m.conv(b + ".upsamplers.0.conv", cout, cout, 3);
    }
    m.norm("conv_norm_out", 64);
    m.conv("conv_out", 3, 64, 3);
    return m;
}
static const Manifest& manifest() { static Manifest m = build_manifest(); return m; }

// ------------------------------------------------------------------ host packing ------
static uint16_t f2bf(float f) {                  // round-to-nearest-even, NaN-safe enough for weights
    uint32_t u; memcpy(&u, &f, 4);
    const uint32_t r = 0x7FFFu + ((u >> 16) & 1u);
    return (uint16_t)((u + r) >> 16);
}

struct DevBuf {                                   // owns one cudaMalloc
    void* p = nullptr;
    ~DevBuf() { if (p) cudaFree(p); }
};
using DevPtr = std::shared_ptr<DevBuf>;
static DevPtr dev_upload(const void* host, size_t bytes) {
    auto b = std::make_shared<DevBuf>();
    SYNT_CUDA(cudaMalloc(&b->p, bytes ? bytes : 4));
    if (bytes) SYNT_CUDA(cudaMemcpy(b->p, host, bytes, cudaMemcpyHostToDevice));
    return b;
}
static DevPtr dev_alloc(size_t bytes) {
    auto b = std::make_shared<DevBuf>();
    SYNT_CUDA(cudaMalloc(&b->p, bytes ? bytes : 4));
    SYNT_CUDA(cudaMemset(b->p, 0, bytes ? bytes : 4));
    return b;
}

// conv weight [Cout][Cin][k][k] (+ optional 1x1 shortcut [Cout][Csc]) -> K-major [Cout][k*k*Cin + Csc]
static std::vector<float> pack_conv(const float* w, int cout, int cin, int k, const float* wsc, int csc) {
    const int taps = k * k, ktot = taps * cin + csc;
    std::vector<float> o((size_t)cout * ktot);
    for (int n = 0; n < cout; ++n) {
        float* row = o.data() + (size_t)n * ktot;
        for (int t = 0; t < taps; ++t)
            for (int c = 0; c < cin; ++c) row[t * cin + c] = w[((size_t)n * cin + c) * taps + t];
        for (int c = 0; c < csc; ++c) row[taps * cin + c] = wsc[(size_t)n * csc + c];
    }
    return o;
}
// Upsample2D (nearest 2x) followed by conv3x3(pad 1) == four 2x2 convolutions on the LOW-res input, one per
// output sub-pixel phase (py, px): output (2y+py, 2x+px) reads low-res rows y+py-1, y+py and columns x+px-1,
// x+px; the 3x3 taps that land on the same low-res pixel are summed (in fp32, before the bf16 rounding).
// [Cout][Cin][3][3] -> phase-stacked K-major [4*Cout][4*Cin]: row = phase*Cout + n, col = (ty*2+tx)*Cin + c.
static std::vector<float> pack_upsample_phases(const float* w, int cout, int cin) {
    std::vector<float> o((size_t)16 * cout * cin, 0.f);
    // low-res tap t of phase p collects 3x3 taps k with (p + k - 1) >> 1 == p - 1 + t (floor division)
    auto lowtap = [](int p, int k) { const int u = p + k - 1; return (u < 0 ? -1 : u >> 1) - (p - 1); };
    for (int py = 0; py < 2; ++py)
        for (int px = 0; px < 2; ++px)
            for (int n = 0; n < cout; ++n) {
                float* row = o.data() + ((size_t)(py * 2 + px) * cout + n) * 4 * cin;
                for (int ky = 0; ky < 3; ++ky)
                    for (int kx = 0; kx < 3; ++kx) {
                        const int t = lowtap(py, ky) * 2 + lowtap(px, kx);
                        for (int c = 0; c < cin; ++c) row[t * cin + c] += w[(((size_t)n * cin + c) * 3 + ky) * 3 + kx];
                    }
            }
    return o;
}
struct WeightDev {                                // one GEMM operand on the device
    DevPtr f32, b16;
    const void* get(bool tc) const { return tc ? b16->p : f32->p; }
};
static WeightDev upload_weight(const std::vector<float>& w, bool want_f32, bool want_bf16) {
    WeightDev d;
    if (want_f32) d.f32 = dev_upload(w.data(), w.size() * 4);
    if (want_bf16) {
        std::vector<uint16_t> h(w.size());
        for (size_t i = 0; i < w.size(); ++i) h[i] = f2bf(w[i]);
        d.b16 = dev_upload(h.data(), h.size() * 2);
    }
    return d;
}

// ------------------------------------------------------------------ model -------------
struct Act {                                      // NHWC activation
    void* p = nullptr; int B = 0, H = 0, W = 0, C = 0;
    float2* stats = nullptr; int stats_slots = 0;  // per-channel (sum, sumsq) partials [B][slots][C] for GroupNorm
    size_t numel() const { return (size_t)B * H * W * C; }
};

struct ResnetW {
    std::string name; int cin, cout; bool shortcut; int temb_off;
    DevPtr g1, b1n, g2, b2n;                      // GroupNorm affine
    WeightDev w1, w2; DevPtr bias1, bias2;        // bias2 = conv2.bias (+ conv_shortcut.bias)
};
struct AttnW {
    std::string name; int C;
    DevPtr g, b; WeightDev wqkv, wo; DevPtr bqkv, bo;
    WeightDev wqkv_tc; DevPtr bqkv_tc;            // (q*log2(e)/sqrt(8) | k | v) projection for attention_tc
};
struct ConvW { std::string name; int cin, cout; WeightDev w; DevPtr b; WeightDev w_up; };   // w_up: pack_upsample_phases

}  // namespace synt

using namespace synt;

struct synt_unet {
    int dt = DT_BF16;
    bool use_tc = true;                           // tcgen05 convs (bf16 mode) vs fp32-FMA convs
    bool want_f32_w = false;
    bool use_v2 = true;                           // persistent halo-tile kernel for 3x3 stride-1 convs
    bool fuse_gn = true;                          // GroupNorm(+SiLU) applied inside conv_tc2 (no normalised tensor in HBM)
    bool fuse_up = true;                          // Upsample2D folded into its conv (sub-pixel phases)
    ConvInW conv_in_w;
    DevPtr conv_in_taps, conv_in_bias;            // conv_in as nine tcgen05 tap tiles (bf16 mode; SYNT_CONV_IN_TC=0: the FMA kernel)
    ConvOutW conv_out_w;
    DevPtr norm_out_g, norm_out_b;
    DevPtr conv_out_frag;                         // conv_out weights as mma.sync B fragments (bf16 mode)
    std::vector<ResnetW> down_res[4], up_res[4];
    std::vector<AttnW> down_attn[4], up_attn[4];
    std::vector<ConvW> downsample, upsample;      // index = block (3 each)
    ResnetW mid_res[2]; AttnW mid_attn;
    int temb_total = 0;
    DevPtr temb_table;                            // [1000][temb_total]
    DevPtr temb_cur, coef_cur, step_ctr;          // fixed per-step buffers (graph friendly)
    DevPtr timesteps_dev, coef_table_dev; int n_steps = 0;
    const unsigned char* step_mask = nullptr; int noise_shared = 0;   // coalition decoding (synt_unet_set_step_mask)
    const long long* image_keys = nullptr;        // per-image Philox stream ids (synt_unet_set_image_keys)
    std::vector<int> timesteps_host;
    Pool pool;
    Pool pool2;                                   // second chain (dual-stream sampling)
    bool dual = false; cudaStream_t side_stream = nullptr; cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    long long launches = 0;
    // CUDA graph cache for one sampling step
    struct GraphKey {
        float* x; int B; const float* z; float* traj; float* eps; int mb; unsigned long long seed; long long off;
        const unsigned char* mask; int shared; const long long* keys;
        bool operator==(const GraphKey& o) const {
            return x == o.x && B == o.B && z == o.z && traj == o.traj && eps == o.eps && mb == o.mb && seed == o.seed &&
                   off == o.off && mask == o.mask && shared == o.shared && keys == o.keys;
        }
    } gkey{};
    cudaGraphExec_t gexec = nullptr;
    long long graph_nodes = 0;
    cudaStream_t own_stream = nullptr;            // capture needs a non-legacy stream
    cudaEvent_t ev_in = nullptr, ev_out = nullptr;
    // per-kernel profiling (bench.py roofline): events around every launch of one eager step
    struct ProfRec { int cat; cudaEvent_t a, b; double flops; int m, n, k; float ms; };
    std::vector<ProfRec> last_recs;
    bool prof = false;
    std::vector<ProfRec> recs;
    // debug tap
    std::string tap_name; float* tap_out = nullptr; long long tap_cap = 0; int tap_C = 0, tap_H = 0, tap_W = 0; bool tap_hit = false;

    ~synt_unet() {
        if (gexec) cudaGraphExecDestroy(gexec);
        if (own_stream) cudaStreamDestroy(own_stream);
        if (side_stream) cudaStreamDestroy(side_stream);
        if (ev_fork) cudaEventDestroy(ev_fork);
        if (ev_join) cudaEventDestroy(ev_join);
        if (ev_in) cudaEventDestroy(ev_in);
        if (ev_out) cudaEventDestroy(ev_out);
    }
};

namespace synt {

static ResnetW make_resnet(const float* P, const std::string& name, int cin, int cout, int temb_off, bool f32, bool b16) {
    const Manifest& m = manifest();
    ResnetW r; r.name = name; r.cin = cin; r.cout = cout; r.shortcut = cin != cout; r.temb_off = temb_off;
    r.g1 = dev_upload(P + m.find(name + ".norm1.weight"), cin * 4);
    r.b1n = dev_upload(P + m.find(name + ".norm1.bias"), cin * 4);
    r.g2 = dev_upload(P + m.find(name + ".norm2.weight"), cout * 4);
    r.b2n = dev_upload(P + m.find(name + ".norm2.bias"), cout * 4);
    r.w1 = upload_weight(pack_conv(P + m.find(name + ".conv1.weight"), cout, cin, 3, nullptr, 0), f32, b16);
    r.bias1 = dev_upload(P + m.find(name + ".conv1.bias"), cout * 4);
    std::vector<float> b2(P + m.find(name + ".conv2.bias"), P + m.find(name + ".conv2.bias") + cout);
    if (r.shortcut) {
        const float* bs = P + m.find(name + ".conv_shortcut.bias");
        for (int i = 0; i < cout; ++i) b2[i] += bs[i];
        r.w2 = upload_weight(pack_conv(P + m.find(name + ".conv2.weight"), cout, cout, 3,
                                       P + m.find(name + ".conv_shortcut.weight"), cin), f32, b16);
    } else {
        r.w2 = upload_weight(pack_conv(P + m.find(name + ".conv2.weight"), cout, cout, 3, nullptr, 0), f32, b16);
    }
    r.bias2 = dev_upload(b2.data(), cout * 4);
    return r;
}
static AttnW make_attn(const float* P, const std::string& name, int C, bool f32, bool b16) {
    const Manifest& m = manifest();
    AttnW a; a.name = name; a.C = C;
    a.g = dev_upload(P + m.find(name + ".group_norm.weight"), C * 4);
    a.b = dev_upload(P + m.find(name + ".group_norm.bias"), C * 4);
    std::vector<float> wqkv((size_t)3 * C * C), bqkv(3 * C);
    const char* nm[3] = {".to_q", ".to_k", ".to_v"};
    for (int i = 0; i < 3; ++i) {
        memcpy(wqkv.data() + (size_t)i * C * C, P + m.find(name + nm[i] + ".weight"), (size_t)C * C * 4);
        memcpy(bqkv.data() + i * C, P + m.find(name + nm[i] + ".bias"), C * 4);
    }
    a.wqkv = upload_weight(wqkv, f32, b16);
    a.bqkv = dev_upload(bqkv.data(), bqkv.size() * 4);
    if (b16) {
        // attention_tc reads the plain (q | k | v) projection; the softmax scale and the exp->exp2 conversion are
        // folded into the q rows
        const float qscale = 1.4426950408889634f / std::sqrt(8.0f);
        std::vector<float> wp(wqkv), bp(bqkv);
        for (size_t i = 0; i < (size_t)C * C; ++i) wp[i] *= qscale;
        for (int i = 0; i < C; ++i) bp[i] *= qscale;
        a.wqkv_tc = upload_weight(wp, false, true);
        a.bqkv_tc = dev_upload(bp.data(), bp.size() * 4);
    }
    std::vector<float> wo(P + m.find(name + ".to_out.0.weight"), P + m.find(name + ".to_out.0.weight") + (size_t)C * C);
    a.wo = upload_weight(wo, f32, b16);
    a.bo = dev_upload(P + m.find(name + ".to_out.0.bias"), C * 4);
    return a;
}
static ConvW make_conv(const float* P, const std::string& name, int cin, int cout, bool f32, bool b16, bool upsampler = false) {
    const Manifest& m = manifest();
    ConvW c; c.name = name; c.cin = cin; c.cout = cout;
    c.w = upload_weight(pack_conv(P + m.find(name + ".weight"), cout, cin, 3, nullptr, 0), f32, b16);
    if (upsampler && b16) c.w_up = upload_weight(pack_upsample_phases(P + m.find(name + ".weight"), cout, cin), false, true);
    c.b = dev_upload(P + m.find(name + ".bias"), cout * 4);
    return c;
}

static void build_unet(synt_unet* u, const float* P) {
    const Manifest& m = manifest();
    const bool b16 = u->dt == DT_BF16 && u->use_tc;
    const bool f32 = !b16;
    // edge convolutions (by-value constant-bank weights)
    {
        const float* w = P + m.find("conv_in.weight");      // [64][3][3][3]
        for (int n = 0; n < 64; ++n)
            for (int c = 0; c < 3; ++c)
                for (int t = 0; t < 9; ++t) u->conv_in_w.w[t * 3 + c][n] = w[((size_t)n * 3 + c) * 9 + t];
        memcpy(u->conv_in_w.b, P + m.find("conv_in.bias"), 64 * 4);
        const float* wo = P + m.find("conv_out.weight");    // [3][64][3][3]
        for (int n = 0; n < 3; ++n)
            for (int c = 0; c < 64; ++c)
                for (int t = 0; t < 9; ++t) u->conv_out_w.w[t][c][n] = wo[((size_t)n * 64 + c) * 9 + t];
        memcpy(u->conv_out_w.b, P + m.find("conv_out.bias"), 3 * 4);
        const char* cit = getenv("SYNT_CONV_IN_TC");
        if (b16 && !(cit && cit[0] == '0')) {
            std::vector<uint16_t> taps((size_t)conv_in_tc_weight_bytes() / 2);
            conv_in_tc_pack_weights(u->conv_in_w, taps.data());
            u->conv_in_taps = dev_upload(taps.data(), taps.size() * 2);
            u->conv_in_bias = dev_upload(u->conv_in_w.b, 64 * 4);
        }
        if (b16) {                                          // B fragments of mma.sync.m16n8k16 (see kernels.cuh)
            std::vector<uint32_t> fr((size_t)36 * 32 * 2);
            auto wv = [&](int tap, int c, int n) -> uint32_t { return n < 3 ? f2bf(wo[((size_t)n * 64 + c) * 9 + tap]) : 0u; };
            for (int tap = 0; tap < 9; ++tap)
                for (int kc = 0; kc < 4; ++kc)
                    for (int lane = 0; lane < 32; ++lane) {
                        const int g = lane >> 2, t = lane & 3, k0 = kc * 16 + t * 2;
                        uint32_t* o = fr.data() + ((size_t)(tap * 4 + kc) * 32 + lane) * 2;
                        o[0] = wv(tap, k0, g) | (wv(tap, k0 + 1, g) << 16);
                        o[1] = wv(tap, k0 + 8, g) | (wv(tap, k0 + 9, g) << 16);
                    }
            u->conv_out_frag = dev_upload(fr.data(), fr.size() * 4);
        }
        u->norm_out_g = dev_upload(P + m.find("conv_norm_out.weight"), 64 * 4);
        u->norm_out_b = dev_upload(P + m.find("conv_norm_out.bias"), 64 * 4);
    }
    std::vector<float> proj_w, proj_b;            // concatenated time_emb_proj of every resnet
    auto add_res = [&](const std::string& name, int cin, int cout) {
        const int off = (int)proj_b.size();
        const float* w = P + m.find(name + ".time_emb_proj.weight");
        const float* b = P + m.find(name + ".time_emb_proj.bias");
        proj_w.insert(proj_w.end(), w, w + (size_t)cout * kTembDim);
        proj_b.insert(proj_b.end(), b, b + cout);
        return make_resnet(P, name, cin, cout, off, f32, b16);
    };
    int out = 64;
    for (int i = 0; i < 4; ++i) {
        const int inp = out; out = kBlockCh[i];
        const std::string b = "down_blocks." + std::to_string(i);
        for (int j = 0; j < 2; ++j) u->down_res[i].push_back(add_res(b + ".resnets." + std::to_string(j), j == 0 ? inp : out, out));
        if (kDownAttn[i]) for (int j = 0; j < 2; ++j) u->down_attn[i].push_back(make_attn(P, b + ".attentions." + std::to_string(j), out, f32, b16));
        if (i != 3) u->downsample.push_back(make_conv(P, b + ".downsamplers.0.conv", out, out, f32, b16));
    }
    u->mid_res[0] = add_res("mid_block.resnets.0", 256, 256);
    u->mid_attn = make_attn(P, "mid_block.attentions.0", 256, f32, b16);
    u->mid_res[1] = add_res("mid_block.resnets.1", 256, 256);
    for (int i = 0; i < 4; ++i) {
        const std::string b = "up_blocks." + std::to_string(i);
        int cout = 0;
        for (int j = 0; j < 3; ++j) {
            int ch, cs; up_resnet_channels(i, j, ch, cs, cout);
            u->up_res[i].push_back(add_res(b + ".resnets." + std::to_string(j), ch + cs, cout));
        }
        if (kUpAttn[i]) for (int j = 0; j < 3; ++j) u->up_attn[i].push_back(make_attn(P, b + ".attentions." + std::to_string(j), cout, f32, b16));
        if (i != 3) u->upsample.push_back(make_conv(P, b + ".upsamplers.0.conv", cout, cout, f32, b16, true));
    }
    u->temb_total = (int)proj_b.size();

    // time-embedding tables for all 1000 train timesteps (batch independent: same t for all b)
    float freqs[32];
    for (int k = 0; k < 32; ++k) freqs[k] = (float)std::exp(-std::log(10000.0) * (double)k / 32.0);
    DevPtr d_freq = dev_upload(freqs, sizeof(freqs));
    DevPtr w1 = dev_upload(P + m.find("time_embedding.linear_1.weight"), 256 * 64 * 4);
    DevPtr b1 = dev_upload(P + m.find("time_embedding.linear_1.bias"), 256 * 4);
    DevPtr w2 = dev_upload(P + m.find("time_embedding.linear_2.weight"), 256 * 256 * 4);
    DevPtr b2 = dev_upload(P + m.find("time_embedding.linear_2.bias"), 256 * 4);
    DevPtr emb = dev_alloc((size_t)kTrainSteps * 256 * 4);
    DevPtr pw = dev_upload(proj_w.data(), proj_w.size() * 4);
    DevPtr pb = dev_upload(proj_b.data(), proj_b.size() * 4);
    u->temb_table = dev_alloc((size_t)kTrainSteps * u->temb_total * 4);
    time_embed_table((const float*)d_freq->p, (const float*)w1->p, (const float*)b1->p, (const float*)w2->p,
                     (const float*)b2->p, kTrainSteps, (float*)emb->p, 0);
    time_proj_table((const float*)emb->p, (const float*)pw->p, (const float*)pb->p, kTrainSteps, u->temb_total,
                    (float*)u->temb_table->p, 0);
    SYNT_CUDA(cudaDeviceSynchronize());
    u->temb_cur = dev_alloc((size_t)u->temb_total * 4);
    u->coef_cur = dev_alloc(8 * 4);
    u->step_ctr = dev_alloc(4);
}

// ------------------------------------------------------------------ forward -----------
enum ProfCat { PC_CONV_TC = 0, PC_CONV_SIMT, PC_GN_STATS, PC_GN_APPLY, PC_ATTN, PC_UPSAMPLE, PC_CONV_IN, PC_CONV_OUT,
               PC_MISC, PC_COUNT };
struct ProfScope {
    synt_unet* u; cudaStream_t s; int cat; double flops; cudaEvent_t a = nullptr; int m = 0, n = 0, k = 0;
    ProfScope(synt_unet* u_, cudaStream_t s_, int cat_, double flops_ = 0.0, int m_ = 0, int n_ = 0, int k_ = 0)
        : u(u_), s(s_), cat(cat_), flops(flops_), m(m_), n(n_), k(k_) {
        if (u->prof) { cudaEventCreate(&a); cudaEventRecord(a, s); }
    }
    ~ProfScope() {
        if (u->prof) { cudaEvent_t b; cudaEventCreate(&b); cudaEventRecord(b, s); u->recs.push_back({cat, a, b, flops, m, n, k, 0.f}); }
    }
};

struct Fwd {
    synt_unet* u; cudaStream_t s; int B; Pool* pool;
    size_t esz() const { return dtype_size(u->dt); }
    Act make(int H, int W, int C) {
        Act a; a.B = B; a.H = H; a.W = W; a.C = C; a.p = pool->alloc(a.numel() * esz()); return a;
    }
    void drop(Act& a) { pool->release(a.p); pool->release(a.stats); a.p = nullptr; a.stats = nullptr; }
    // per-channel statistics by the stand-alone kernel (tensors whose producer does not fuse them)
    void standalone_stats(Act& a) {
        const int HW = a.H * a.W;
        a.stats_slots = gn_num_chunks(B, HW);
        a.stats = (float2*)pool->alloc((size_t)B * a.stats_slots * a.C * sizeof(float2));
        ProfScope ps(u, s, PC_GN_STATS);
        gn_stats(a.p, a.C, nullptr, 0, u->dt, B, HW, a.C, a.stats, a.stats_slots, s);
        ++u->launches;
    }
    void tap(const std::string& name, const Act& a) {
        if (!u->tap_out || name != u->tap_name) return;
        SYNT_CHECK((long long)a.numel() <= u->tap_cap, "debug tap buffer too small");
        nhwc_to_nchw_f32(a.p, u->dt, a.B, a.H * a.W, a.C, u->tap_out, s);
        u->tap_C = a.C; u->tap_H = a.H; u->tap_W = a.W; u->tap_hit = true;
        ++u->launches;
    }
    // GroupNorm statistics of concat(x0, x1) -> per-(b, c) scale/shift
    float2* gn_scale_shift(const Act& x0, const Act* x1, const DevPtr& gamma, const DevPtr& beta) {
        const int C = x0.C + (x1 ? x1->C : 0), HW = x0.H * x0.W;
        SYNT_CHECK(x0.stats && (!x1 || x1->stats), "GroupNorm input without statistics");
        float2* ss = (float2*)pool->alloc((size_t)B * C * sizeof(float2));
        ProfScope ps(u, s, PC_GN_STATS);
        gn_finalize_channels(x0.stats, x0.stats_slots, x0.C, x1 ? x1->stats : nullptr, x1 ? x1->stats_slots : 0,
                             x1 ? x1->C : 0, B, kGroups, HW, kGnEps, (const float*)gamma->p, (const float*)beta->p, ss, s);
        ++u->launches;
        return ss;
    }
    Act gn_act(const Act& x0, const Act* x1, const DevPtr& gamma, const DevPtr& beta, bool silu) {
        SYNT_CHECK(x0.stats && (!x1 || x1->stats), "GroupNorm input without statistics");
        float2* ss = gn_scale_shift(x0, x1, gamma, beta);
        Act o = make(x0.H, x0.W, x0.C + (x1 ? x1->C : 0));
        ProfScope ps(u, s, PC_GN_APPLY, 0.0, B * x0.H * x0.W, o.C, x1 ? x1->C : 0);
        gn_apply_fused(x0.p, x0.stats, x0.stats_slots, x0.C, x1 ? x1->p : nullptr, x1 ? x1->stats : nullptr,
                       x1 ? x1->stats_slots : 0, x1 ? x1->C : 0, u->dt, B, x0.H * x0.W, kGroups, kGnEps,
                       (const float*)gamma->p, (const float*)beta->p, ss, silu ? 1 : 0, o.p, s);
        pool->release(ss);
        ++u->launches;
        return o;
    }
    bool fused_ok(const ConvArgs& a) const { return u->dt == DT_BF16 && u->use_tc && u->use_v2 && conv_tc2_supported(a); }
    // `out` receives GroupNorm statistics when want_stats: fused in the conv_tc2 epilogue, else stand-alone
    void conv(ConvArgs& a, const WeightDev& w, Act* out = nullptr, bool want_stats = false) {
        const bool v2 = fused_ok(a);
        const bool tc = v2 || (u->dt == DT_BF16 && u->use_tc && conv_tc_supported(a));
        a.weight = w.get(tc);
        if (want_stats && v2) {
            out->stats_slots = conv_tc2_stats_slots(a);
            out->stats = (float2*)pool->alloc((size_t)B * out->stats_slots * a.Cout * sizeof(float2));
            a.stats_out = out->stats;
        }
        {
            // up2x: algorithmic FLOPs of the reference's conv3x3 on the upsampled image (executed: 4/9 of that)
            const int up = a.up2x ? 4 : 1;
            ProfScope ps(u, s, tc ? PC_CONV_TC : PC_CONV_SIMT, 2.0 * B * a.Ho * a.Wo * up * (double)a.Cout * a.ktot(),
                         B * a.Ho * a.Wo * up, a.Cout, a.ktot());
            if (v2) conv_tc2(a, s);
            else if (tc) conv_tc(a, s);
            else conv_simt(a, u->dt, s);
            ++u->launches;
        }
        if (want_stats && !v2) standalone_stats(*out);
    }
    Act resnet(const ResnetW& r, const Act& x0, const Act* x1) {
        const int H = x0.H, W = x0.W;
        // probe: can both convs take their GroupNorm'd input through the in-kernel transform?
        ConvArgs probe; probe.B = B; probe.H = H; probe.W = W; probe.Ho = H; probe.Wo = W;
        probe.Cin = x0.C; probe.Cin1 = x1 ? x1->C : 0; probe.Cout = r.cout;
        const bool fuse = u->fuse_gn && fused_ok(probe);
        Act h1 = make(H, W, r.cout);
        {
            ConvArgs c; c.B = B; c.H = H; c.W = W; c.Ho = H; c.Wo = W; c.Cout = r.cout;
            c.bias = (const float*)r.bias1->p; c.bias2 = (const float*)u->temb_cur->p + r.temb_off; c.out = h1.p;
            if (fuse) {
                float2* ss = gn_scale_shift(x0, x1, r.g1, r.b1n);
                c.in = x0.p; c.Cin = x0.C;
                if (x1) { c.in1 = x1->p; c.Cin1 = x1->C; }
                c.gn_ss = ss; c.gn_mode = 2;
                conv(c, r.w1, &h1, true);
                pool->release(ss);
            } else {
                Act a = gn_act(x0, x1, r.g1, r.b1n, true);
                c.in = a.p; c.Cin = r.cin;
                conv(c, r.w1, &h1, true);
                drop(a);
            }
        }
        Act o = make(H, W, r.cout);
        {
            ConvArgs c; c.B = B; c.H = H; c.W = W; c.Cin = r.cout; c.Ho = H; c.Wo = W; c.Cout = r.cout;
            c.bias = (const float*)r.bias2->p; c.out = o.p;
            if (r.shortcut) {
                c.sc0 = x0.p; c.sc0_C = x0.C;
                if (x1) { c.sc1 = x1->p; c.sc1_C = x1->C; }
            } else {
                SYNT_CHECK(x1 == nullptr, "identity residual with a concatenated input");
                c.residual = x0.p;
            }
            if (fuse) {
                float2* ss = gn_scale_shift(h1, nullptr, r.g2, r.b2n);
                c.in = h1.p; c.gn_ss = ss; c.gn_mode = 2;
                conv(c, r.w2, &o, true);
                pool->release(ss);
            } else {
                Act a2 = gn_act(h1, nullptr, r.g2, r.b2n, true);
                c.in = a2.p;
                conv(c, r.w2, &o, true);
                drop(a2);
            }
        }
        drop(h1);
        tap(r.name, o);
        return o;
    }
    Act attention(const AttnW& w, const Act& x) {
        const int H = x.H, W = x.W, C = w.C;
        const bool tc = u->dt == DT_BF16 && u->use_tc && attention_tc_supported(H * W, C);
        Act qkv = make(H, W, 3 * C);
        {
            ConvArgs c; c.B = B; c.H = H; c.W = W; c.Cin = C; c.KH = c.KW = 1; c.pad = 0; c.Ho = H; c.Wo = W;
            c.Cout = qkv.C; c.bias = (const float*)(tc ? w.bqkv_tc : w.bqkv)->p; c.out = qkv.p;
            // (the halo geometry would re-transform a 46 KB tile per N tile; the 1x1 variant's affine-only pass over 32 KB is cheap)
            if (u->fuse_gn && fused_ok(c) && (c.Cout <= 256 || conv_tc2_is_k1(c))) {
                float2* ss = gn_scale_shift(x, nullptr, w.g, w.b);
                c.in = x.p; c.gn_ss = ss; c.gn_mode = 1;
                conv(c, tc ? w.wqkv_tc : w.wqkv);
                pool->release(ss);
            } else {
                Act a = gn_act(x, nullptr, w.g, w.b, false);
                c.in = a.p;
                conv(c, tc ? w.wqkv_tc : w.wqkv);
                drop(a);
            }
        }
        Act o = make(H, W, C);
        {
            ProfScope ps(u, s, PC_ATTN, 4.0 * B * (double)(H * W) * (H * W) * C);
            if (tc) {
                attention_tc(qkv.p, B, H * W, C, nullptr, o.p, s);     // V^T is built inside the kernel
            } else {
                attention_simt(qkv.p, u->dt, B, H * W, C, o.p, s);
            }
        }
        ++u->launches;
        drop(qkv);
        Act out = make(H, W, C);
        {
            ConvArgs c; c.in = o.p; c.B = B; c.H = H; c.W = W; c.Cin = C; c.KH = c.KW = 1; c.pad = 0; c.Ho = H; c.Wo = W;
            c.Cout = C; c.bias = (const float*)w.bo->p; c.residual = x.p; c.out = out.p;
            conv(c, w.wo, &out, true);
        }
        drop(o);
        tap(w.name, out);
        return out;
    }
    Act down(const ConvW& w, const Act& x) {
        Act o = make(x.H / 2, x.W / 2, w.cout);
        ConvArgs c; c.in = x.p; c.B = B; c.H = x.H; c.W = x.W; c.Cin = w.cin; c.stride = 2; c.Ho = o.H; c.Wo = o.W;
        c.Cout = w.cout; c.bias = (const float*)w.b->p; c.out = o.p;
        conv(c, w.w, &o, true);
        tap(w.name.substr(0, w.name.size() - 5), o);           // strip ".conv"
        return o;
    }
    Act up(const ConvW& w, const Act& x) {
        if (u->fuse_up && w.w_up.b16) {
            // nearest-2x upsample folded into the conv: four sub-pixel 2x2 convolutions on the low-res input
            ConvArgs c; c.in = x.p; c.B = B; c.H = x.H; c.W = x.W; c.Cin = w.cin; c.Ho = x.H; c.Wo = x.W;
            c.Cout = w.cout; c.bias = (const float*)w.b->p; c.up2x = 1;
            if (fused_ok(c)) {
                Act o = make(x.H * 2, x.W * 2, w.cout);
                c.out = o.p;
                conv(c, w.w_up, &o, true);
                tap(w.name.substr(0, w.name.size() - 5), o);
                return o;
            }
        }
        Act up2 = make(x.H * 2, x.W * 2, x.C);
        { ProfScope ps(u, s, PC_UPSAMPLE); upsample_nearest2x(x.p, u->dt, B, x.H, x.W, x.C, up2.p, s); }
        ++u->launches;
        Act o = make(up2.H, up2.W, w.cout);
        ConvArgs c; c.in = up2.p; c.B = B; c.H = up2.H; c.W = up2.W; c.Cin = w.cin; c.Ho = o.H; c.Wo = o.W;
        c.Cout = w.cout; c.bias = (const float*)w.b->p; c.out = o.p;
        conv(c, w.w, &o, true);
        drop(up2);
        tap(w.name.substr(0, w.name.size() - 5), o);
        return o;
    }

    // eps = UNet(x, t) for one micro-batch; temb_cur must already hold the row of t.
    void run(const float* x_nchw, float* eps_nchw, const SchedArgs& sch) {
        Pool::Scope scope(*pool);                                   // an exception below returns every block still out
        Act h = make(kImg, kImg, 64);
        const int in_slots = conv_in3_stats_slots(kImg, kImg, u->dt);
        if (in_slots > 0) {                                     // statistics of the first GroupNorm fused into conv_in
            h.stats_slots = in_slots;
            h.stats = (float2*)pool->alloc((size_t)B * in_slots * 64 * sizeof(float2));
        }
        {
            ProfScope ps(u, s, PC_CONV_IN, 2.0 * B * kImg * kImg * 64.0 * 27);
            if (u->conv_in_taps && in_slots == 16)
                conv_in_tc(x_nchw, B, u->conv_in_taps->p, (const float*)u->conv_in_bias->p, h.p, h.stats, s);
            else
                conv_in3(x_nchw, u->conv_in_w, B, kImg, kImg, h.p, u->dt, h.stats, s);
        }
        ++u->launches;
        if (in_slots == 0) standalone_stats(h);
        tap("conv_in", h);
        std::vector<Act> skips; skips.push_back(h);
        Act cur = h;                                            // `cur` aliases the top skip
        for (int i = 0; i < 4; ++i) {
            for (int j = 0; j < 2; ++j) {
                Act r = resnet(u->down_res[i][j], cur, nullptr);
                if (kDownAttn[i]) { Act a = attention(u->down_attn[i][j], r); drop(r); r = a; }
                skips.push_back(r); cur = r;
            }
            if (i != 3) { Act d = down(u->downsample[i], cur); skips.push_back(d); cur = d; }
        }
        // mid block (cur is still owned by the skip stack)
        Act m0 = resnet(u->mid_res[0], cur, nullptr);
        Act m1 = attention(u->mid_attn, m0); drop(m0);
        Act hcur = resnet(u->mid_res[1], m1, nullptr); drop(m1);
        for (int i = 0; i < 4; ++i) {
            for (int j = 0; j < 3; ++j) {
                Act skip = skips.back(); skips.pop_back();
                Act r = resnet(u->up_res[i][j], hcur, &skip);
                drop(hcur); drop(skip);
                if (kUpAttn[i]) { Act a = attention(u->up_attn[i][j], r); drop(r); r = a; }
                hcur = r;
            }
            if (i != 3) { Act o = up(u->upsample[i], hcur); drop(hcur); hcur = o; }
        }
        float2* ss = gn_scale_shift(hcur, nullptr, u->norm_out_g, u->norm_out_b);
        { ProfScope ps(u, s, PC_CONV_OUT, 2.0 * B * kImg * kImg * 3.0 * 576); conv_out3(hcur.p, u->dt, ss, u->conv_out_w, u->conv_out_frag ? u->conv_out_frag->p : nullptr, B, kImg, kImg, eps_nchw, sch, s); }
        ++u->launches;
        pool->release(ss);
        drop(hcur);
        scope.commit();
    }
};

static int default_micro_batch(int B) { return B <= 64 ? B : 64; }     // measured: no L2 benefit from smaller slices

// one sampling step for all micro-batches (captured into the CUDA graph)
static void sample_step(synt_unet* u, float* x, int B, const float* z, unsigned long long seed, long long image_offset,
                        float* traj, float* eps_tap, int mb, cudaStream_t s) {
    const size_t img = (size_t)3 * kImg * kImg;
    select_timestep((const float*)u->temb_table->p, u->temb_total, (const float*)u->coef_table_dev->p,
                    (const int*)u->timesteps_dev->p, (const int*)u->step_ctr->p, 0, (float*)u->temb_cur->p,
                    (float*)u->coef_cur->p, s);
    ++u->launches;
    auto chain = [&](int b0, int nb, cudaStream_t cs, Pool* pool) {
        Fwd f{u, cs, nb, pool};
        SchedArgs sch;
        sch.x = x + b0 * img; sch.coef = (const float*)u->coef_cur->p;
        sch.z = z ? z + b0 * img : nullptr; sch.z_step_stride = (long long)B * img;
        sch.seed = seed; sch.step_ptr = (const int*)u->step_ctr->p;
        sch.traj = traj ? traj + b0 * img : nullptr; sch.traj_step_stride = (long long)B * img;
        sch.eps_step_stride = (long long)B * img;
        sch.image_offset = u->noise_shared ? image_offset : image_offset + b0;
        sch.step_mask = u->step_mask ? u->step_mask + b0 : nullptr; sch.mask_stride = B; sch.noise_shared = u->noise_shared;
        sch.image_keys = u->image_keys ? u->image_keys + b0 : nullptr;
        f.run(x + b0 * img, eps_tap ? eps_tap + b0 * img : nullptr, sch);
    };
    if (u->dual && B >= 2 && mb >= B) {
        // two independent half-batch chains on two streams: the tails / small kernels of one chain are
        // filled by the other (captured as two parallel branches of the step graph)
        const int h0 = B / 2;
        SYNT_CUDA(cudaEventRecord(u->ev_fork, s));
        SYNT_CUDA(cudaStreamWaitEvent(u->side_stream, u->ev_fork, 0));
        chain(0, h0, s, &u->pool);
        chain(h0, B - h0, u->side_stream, &u->pool2);
        SYNT_CUDA(cudaEventRecord(u->ev_join, u->side_stream));
        SYNT_CUDA(cudaStreamWaitEvent(s, u->ev_join, 0));
    } else {
        for (int b0 = 0; b0 < B; b0 += mb) chain(b0, B - b0 < mb ? B - b0 : mb, s, &u->pool);
    }
    advance_step((int*)u->step_ctr->p, s);
    ++u->launches;
}

}  // namespace synt

// ======================================================================= C ABI ========
#define SYNT_TRY try {
#define SYNT_CATCH                                                                    \
    } catch (const synt::Error& e) { synt::g_last_error = e.what(); return e.code;    \
    } catch (const std::exception& e) { synt::g_last_error = e.what(); return -1; }   \
    return 0;

extern "C" {

const char* synt_version(void) { return "synt_isic_b200 0.1 (sm_100a)"; }
const char* synt_last_error(void) { return synt::g_last_error.c_str(); }

int synt_unet_num_params(void) { return (int)manifest().params.size(); }
long long synt_unet_total_param_count(void) { return manifest().total; }
int synt_unet_param_info(int i, char* name, int cap, long long* numel, long long* offset) {
    SYNT_TRY
    SYNT_CHECK(i >= 0 && i < (int)manifest().params.size(), "param index out of range");
    const ParamSpec& p = manifest().params[i];
    if (name && cap > 0) { strncpy(name, p.name.c_str(), cap - 1); name[cap - 1] = 0; }
    if (numel) *numel = p.numel;
    if (offset) *offset = p.offset;
    SYNT_CATCH
}

int synt_unet_create(const float* params_host, long long n_params, int dtype, synt_unet_t** out) {
    SYNT_TRY
    SYNT_CHECK(params_host && out, "null argument");
    SYNT_CHECK(n_params == manifest().total, "parameter blob has the wrong length (expected 25,304,963 floats)");
    SYNT_CHECK(dtype == DT_F32 || dtype == DT_BF16, "dtype must be SYNT_DTYPE_F32 or SYNT_DTYPE_BF16");
    int ndev = 0;
    SYNT_CUDA(cudaGetDeviceCount(&ndev));
    SYNT_CHECK(ndev > 0, "no CUDA device: this library has no CPU fallback");
    cudaDeviceProp prop; int dev = 0;
    SYNT_CUDA(cudaGetDevice(&dev));
    SYNT_CUDA(cudaGetDeviceProperties(&prop, dev));
    SYNT_CHECK(prop.major == 10, "synt_isic_b200 is built for sm_100a (Blackwell B200) only");
    std::unique_ptr<synt_unet> u(new synt_unet());
    u->dt = dtype;
    const char* force = getenv("SYNT_FORCE_SIMT");
    u->use_tc = dtype == DT_BF16 && !(force && force[0] == '1');
    const char* v2 = getenv("SYNT_CONV_V2");
    u->use_v2 = !(v2 && v2[0] == '0');
    const char* du = getenv("SYNT_DUAL");
    u->dual = du && du[0] == '1';
    SYNT_CUDA(cudaStreamCreateWithFlags(&u->side_stream, cudaStreamNonBlocking));
    SYNT_CUDA(cudaEventCreateWithFlags(&u->ev_fork, cudaEventDisableTiming));
    SYNT_CUDA(cudaEventCreateWithFlags(&u->ev_join, cudaEventDisableTiming));
    const char* fg = getenv("SYNT_FUSE_GN");
    u->fuse_gn = !(fg && fg[0] == '0');
    const char* fu = getenv("SYNT_FUSE_UP");
    u->fuse_up = !(fu && fu[0] == '0');
    SYNT_CUDA(cudaStreamCreateWithFlags(&u->own_stream, cudaStreamNonBlocking));
    SYNT_CUDA(cudaEventCreateWithFlags(&u->ev_in, cudaEventDisableTiming));
    SYNT_CUDA(cudaEventCreateWithFlags(&u->ev_out, cudaEventDisableTiming));
    build_unet(u.get(), params_host);
    *out = u.release();
    SYNT_CATCH
}
int synt_unet_destroy(synt_unet_t* h) { delete h; return 0; }

int synt_unet_forward(synt_unet_t* h, const float* x, int B, int t, float* eps, void* stream) {
    return synt_unet_debug_forward(h, x, B, t, nullptr, eps, 0, nullptr, nullptr, nullptr, stream);
}

int synt_unet_debug_forward(synt_unet_t* h, const float* x, int B, int t, const char* tap, float* out, long long cap,
                            int* C, int* H, int* W, void* stream) {
    SYNT_TRY
    SYNT_CHECK(h && x && B > 0, "bad argument");
    SYNT_CHECK(t >= 0 && t < kTrainSteps, "timestep out of range [0, 1000)");
    cudaStream_t s = (cudaStream_t)stream;
    const size_t img = (size_t)3 * kImg * kImg;
    const bool dbg = tap != nullptr;
    float* eps = dbg ? nullptr : out;
    if (dbg) { h->tap_name = tap; h->tap_out = out; h->tap_cap = cap; h->tap_hit = false; }
    select_timestep((const float*)h->temb_table->p, h->temb_total, nullptr, nullptr, nullptr, t,
                    (float*)h->temb_cur->p, nullptr, s);
    ++h->launches;
    const int mb = default_micro_batch(B);
    float* dbg_eps = nullptr;
    if (dbg && std::string(tap) == "conv_out") { dbg_eps = out; }
    for (int b0 = 0; b0 < B; b0 += mb) {
        const int nb = B - b0 < mb ? B - b0 : mb;
        SYNT_CHECK(!dbg || nb == B, "debug taps need B <= 16");
        Fwd f{h, s, nb, &h->pool};
        SchedArgs sch;
        f.run(x + b0 * img, dbg ? dbg_eps : eps + b0 * img, sch);
    }
    if (dbg) {
        h->tap_out = nullptr;
        if (dbg_eps) { h->tap_C = 3; h->tap_H = kImg; h->tap_W = kImg; h->tap_hit = true; }
        SYNT_CHECK(h->tap_hit, std::string("unknown debug tap: ") + tap);
        if (C) *C = h->tap_C; if (H) *H = h->tap_H; if (W) *W = h->tap_W;
    }
    SYNT_CATCH
}

int synt_unet_set_schedule(synt_unet_t* h, int n_steps, const int* timesteps, const float* coef) {
    SYNT_TRY
    SYNT_CHECK(h && timesteps && coef && n_steps > 0 && n_steps <= kTrainSteps, "bad schedule");
    for (int i = 0; i < n_steps; ++i) SYNT_CHECK(timesteps[i] >= 0 && timesteps[i] < kTrainSteps, "timestep out of range");
    h->timesteps_dev = dev_upload(timesteps, (size_t)n_steps * 4);
    h->coef_table_dev = dev_upload(coef, (size_t)n_steps * 5 * 4);
    h->timesteps_host.assign(timesteps, timesteps + n_steps);
    h->n_steps = n_steps;
    if (h->gexec) { cudaGraphExecDestroy(h->gexec); h->gexec = nullptr; }
    SYNT_CATCH
}

int synt_unet_sample(synt_unet_t* h, float* x, int B, const float* z, unsigned long long seed, long long image_offset,
                     float* traj, float* eps_tap, int step_begin, int step_end, int micro_batch, int use_graph,
                     void* stream) {
    SYNT_TRY
    SYNT_CHECK(h && x && B > 0, "bad argument");
    SYNT_CHECK(h->n_steps > 0, "synt_unet_set_schedule must be called first");
    SYNT_CHECK(step_begin >= 0 && step_begin <= step_end && step_end <= h->n_steps, "bad step range");
    cudaStream_t caller = (cudaStream_t)stream;
    // the legacy default stream cannot be captured: hop onto the handle's own stream
    const bool hop = use_graph && (caller == nullptr || caller == cudaStreamLegacy || caller == cudaStreamPerThread);
    cudaStream_t s = hop ? h->own_stream : caller;
    struct Rejoin {                                         // make the caller's stream wait for us on every exit
        synt_unet* h; cudaStream_t from, to; bool on;
        ~Rejoin() { if (on) { cudaEventRecord(h->ev_out, from); cudaStreamWaitEvent(to, h->ev_out, 0); } }
    } rejoin{h, s, caller, hop};
    if (hop) {
        SYNT_CUDA(cudaEventRecord(h->ev_in, caller));
        SYNT_CUDA(cudaStreamWaitEvent(s, h->ev_in, 0));
    }
    const int mb = micro_batch > 0 ? (micro_batch < B ? micro_batch : B) : default_micro_batch(B);
    const int n = step_end - step_begin;
    if (n == 0) return 0;
    set_step((int*)h->step_ctr->p, step_begin, s);          // kernel argument: no host buffer to outlive, no host sync
    ++h->launches;
    if (!use_graph) {
        for (int i = 0; i < n; ++i) sample_step(h, x, B, z, seed, image_offset, traj, eps_tap, mb, s);
        return 0;
    }
    synt_unet::GraphKey key{x, B, z, traj, eps_tap, mb, seed, image_offset, h->step_mask, h->noise_shared, h->image_keys};
    int done = 0;
    if (!h->gexec || !(key == h->gkey)) {
        if (h->gexec) { cudaGraphExecDestroy(h->gexec); h->gexec = nullptr; }
        // one eager step first: sizes the pool (no cudaMalloc may happen during capture)
        sample_step(h, x, B, z, seed, image_offset, traj, eps_tap, mb, s);
        done = 1;
        if (n > 1) {
            cudaGraph_t g = nullptr;
            SYNT_CUDA(cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal));
            try {
                sample_step(h, x, B, z, seed, image_offset, traj, eps_tap, mb, s);
            } catch (...) {
                cudaStreamEndCapture(s, &g);
                if (g) cudaGraphDestroy(g);
                throw;
            }
            SYNT_CUDA(cudaStreamEndCapture(s, &g));
            size_t nn = 0;
            SYNT_CUDA(cudaGraphGetNodes(g, nullptr, &nn));
            h->graph_nodes = (long long)nn;
            SYNT_CUDA(cudaGraphInstantiate(&h->gexec, g, 0));
            SYNT_CUDA(cudaGraphDestroy(g));
            h->gkey = key;
            h->launches -= h->graph_nodes;                  // capture did not execute anything
        }
    }
    for (int i = done; i < n; ++i) {
        SYNT_CUDA(cudaGraphLaunch(h->gexec, s));
        h->launches += h->graph_nodes;
    }
    SYNT_CATCH
}

int synt_unet_set_step_mask(synt_unet_t* h, const unsigned char* mask_dev, int noise_shared) {
    SYNT_TRY
    SYNT_CHECK(h, "bad argument");
    h->step_mask = mask_dev; h->noise_shared = noise_shared ? 1 : 0;
    SYNT_CATCH
}

int synt_unet_set_image_keys(synt_unet_t* h, const long long* keys_dev) {
    SYNT_TRY
    SYNT_CHECK(h, "bad argument");
    h->image_keys = keys_dev;
    SYNT_CATCH
}

int synt_unet_generate_host(synt_unet_t* h, const float* xT_host, int B, unsigned long long seed, long long image_offset,
                            int micro_batch, unsigned char* images_u8_host, float* x_final_host) {
    SYNT_TRY
    SYNT_CHECK(h && xT_host && B > 0, "bad argument");
    const size_t n = (size_t)B * 3 * kImg * kImg;
    float* x = (float*)h->pool.alloc(n * 4);
    unsigned char* u8 = (unsigned char*)h->pool.alloc(n);
    cudaStream_t s = h->own_stream;
    SYNT_CUDA(cudaMemcpyAsync(x, xT_host, n * 4, cudaMemcpyHostToDevice, s));
    int rc = synt_unet_sample(h, x, B, nullptr, seed, image_offset, nullptr, nullptr, 0, h->n_steps, micro_batch, 1, s);
    if (rc != 0) { h->pool.release(x); h->pool.release(u8); return rc; }
    if (images_u8_host) {
        to_uint8_hwc(x, B, kImg, kImg, 0, u8, s);
        ++h->launches;
        SYNT_CUDA(cudaMemcpyAsync(images_u8_host, u8, n, cudaMemcpyDeviceToHost, s));
    }
    if (x_final_host) SYNT_CUDA(cudaMemcpyAsync(x_final_host, x, n * 4, cudaMemcpyDeviceToHost, s));
    SYNT_CUDA(cudaStreamSynchronize(s));
    h->pool.release(x); h->pool.release(u8);
    SYNT_CATCH
}

__global__ void synt_spin_kernel(long long cycles) {
    const long long t0 = clock64();
    while (clock64() - t0 < cycles) { }
}

int synt_unet_profile_step(synt_unet_t* h, float* x, int B, int micro_batch, double* ms_out, double* flops_out,
                           int* launches_out, void* stream) {
    SYNT_TRY
    SYNT_CHECK(h && x && B > 0 && ms_out && flops_out && launches_out, "bad argument");
    SYNT_CHECK(h->n_steps > 0, "synt_unet_set_schedule must be called first");
    cudaStream_t s = (cudaStream_t)stream;
    const int mb = micro_batch > 0 ? (micro_batch < B ? micro_batch : B) : default_micro_batch(B);
    int zero = 0;
    SYNT_CUDA(cudaMemcpyAsync(h->step_ctr->p, &zero, 4, cudaMemcpyHostToDevice, s));
    SYNT_CUDA(cudaStreamSynchronize(s));
    sample_step(h, x, B, nullptr, 1, 0, nullptr, nullptr, mb, s);          // warm: pool sized, attributes set
    SYNT_CUDA(cudaStreamSynchronize(s));
    // a spin kernel lets the host enqueue the whole step ahead of the GPU, so that the event pairs
    // bracket kernel execution only (no host launch gaps inside a pair)
    synt_spin_kernel<<<1, 1, 0, s>>>(60000000ll);
    h->prof = true; h->recs.clear();
    try { sample_step(h, x, B, nullptr, 1, 0, nullptr, nullptr, mb, s); } catch (...) { h->prof = false; throw; }
    h->prof = false;
    SYNT_CUDA(cudaStreamSynchronize(s));
    for (int i = 0; i < PC_COUNT; ++i) { ms_out[i] = 0; flops_out[i] = 0; launches_out[i] = 0; }
    for (auto& r : h->recs) {
        float ms = 0.f;
        SYNT_CUDA(cudaEventElapsedTime(&ms, r.a, r.b));
        ms_out[r.cat] += ms; flops_out[r.cat] += r.flops; launches_out[r.cat] += 1;
        cudaEventDestroy(r.a); cudaEventDestroy(r.b);
        r.ms = ms;
    }
    h->last_recs = h->recs;
    h->recs.clear();
    SYNT_CATCH
}

// per-launch records of the last synt_unet_profile_step: rows of {category, ms, flops, M, N, K}
int synt_unet_profile_records(synt_unet_t* h, double* rows6, int cap) {
    if (!h) return 0;
    int n = 0;
    for (auto& r : h->last_recs) {
        if (n >= cap) break;
        double* o = rows6 + (size_t)n * 6;
        o[0] = r.cat; o[1] = r.ms; o[2] = r.flops; o[3] = r.m; o[4] = r.n; o[5] = r.k;
        ++n;
    }
    return n;
}

long long synt_unet_workspace_bytes(synt_unet_t* h) { return h ? (long long)h->pool.total_bytes() : 0; }
long long synt_unet_launch_count(synt_unet_t* h) { return h ? h->launches : 0; }

// ---------------------------------------------------------------- scheduler ----------
int synt_ddpm_tables(int num_train, int schedule, float beta_start, float beta_end, int n_steps, long long* timesteps_out,
                     float* coef_out, float* acp_out) {
    SYNT_TRY
    SYNT_CHECK(num_train > 0 && n_steps > 0 && n_steps <= num_train, "bad step counts");
    std::vector<float> betas(num_train), acp(num_train);
    if (schedule == 0) {                                    // betas_for_alpha_bar: float64 maths, one cast
        auto abar = [](double t) { double c = std::cos((t + 0.008) / 1.008 * M_PI / 2); return c * c; };
        for (int i = 0; i < num_train; ++i) {
            const double t1 = (double)i / num_train, t2 = (double)(i + 1) / num_train;
            double b = 1.0 - abar(t2) / abar(t1);
            betas[i] = (float)(b < 0.999 ? b : 0.999);
        }
    } else if (schedule == 1) {                             // torch.linspace(beta_start, beta_end, n, float32)
        const float step = (beta_end - beta_start) / (float)(num_train - 1);
        for (int i = 0; i < num_train; ++i)
            betas[i] = i < num_train / 2 ? beta_start + step * (float)i : beta_end - step * (float)(num_train - i - 1);
    } else {
        throw Error(-3, "unknown beta schedule");
    }
    // torch.cumprod on a CPU float32 tensor accumulates in double (at::acc_type<float,false>)
    double run = 1.0;
    for (int i = 0; i < num_train; ++i) { const float a = 1.0f - betas[i]; run *= (double)a; acp[i] = (float)run; }
    if (acp_out) memcpy(acp_out, acp.data(), (size_t)num_train * 4);
    const int ratio = num_train / n_steps;                  // "leading" spacing, steps_offset 0
    for (int i = 0; i < n_steps; ++i) {
        const long long t = (long long)(n_steps - 1 - i) * ratio;
        const long long prev = i == n_steps - 1 ? -1 : (long long)(n_steps - 2 - i) * ratio;
        if (timesteps_out) timesteps_out[i] = t;
        if (coef_out) {
            const float a_t = acp[t], a_prev = prev >= 0 ? acp[prev] : 1.0f;
            const float b_t = 1.0f - a_t, b_prev = 1.0f - a_prev;
            const float cur_a = a_t / a_prev, cur_b = 1.0f - cur_a;
            float var = (1.0f - a_prev) / (1.0f - a_t) * cur_b;
            if (var < 1e-20f) var = 1e-20f;
            float* c = coef_out + (size_t)i * 5;
            c[0] = sqrtf(b_t); c[1] = sqrtf(a_t);
            c[2] = (sqrtf(a_prev) * cur_b) / b_t;
            c[3] = sqrtf(cur_a) * b_prev / b_t;
            c[4] = t > 0 ? sqrtf(var) : 0.0f;
        }
    }
    SYNT_CATCH
}

int synt_ddpm_step(const float* eps, const float* x, const float* z, float* out, long long n, const float* c, void* stream) {
    SYNT_TRY
    SYNT_CHECK(eps && x && out && c && n > 0, "bad argument");
    ddpm_step(eps, x, (c[4] != 0.f) ? z : nullptr, out, n, c[0], c[1], c[2], c[3], c[4], (cudaStream_t)stream);
    SYNT_CATCH
}

int synt_to_uint8(const float* x, int B, int H, int W, int mode, unsigned char* out, void* stream) {
    SYNT_TRY
    SYNT_CHECK(x && out && B > 0, "bad argument");
    to_uint8_hwc(x, B, H, W, mode, out, (cudaStream_t)stream);
    SYNT_CATCH
}

}  // extern "C"

// ---------------------------------------------------------------- test hook ----------
// One convolution through either GEMM carrier, on caller-provided NHWC tensors (kernel-level
// parity tests: tcgen05 vs fp32-FMA vs torch.nn.functional.conv2d).
extern "C" int synt_debug_conv(int use_tc, int act_dtype, const void* in, int B, int H, int W, int Cin, int K,
                               int stride, int pad, const void* sc0, int sc0_C, const void* sc1, int sc1_C,
                               int sc_stride, const void* weight, const float* bias, const float* bias2,
                               const void* residual, int relu, void* out, int Cout, void* stream) {
    SYNT_TRY
    ConvArgs a;
    a.in = in; a.B = B; a.H = H; a.W = W; a.Cin = Cin; a.KH = a.KW = K; a.stride = stride; a.pad = pad;
    a.Ho = (H + 2 * pad - K) / stride + 1; a.Wo = (W + 2 * pad - K) / stride + 1; a.Cout = Cout;
    a.sc0 = sc0; a.sc0_C = sc0_C; a.sc1 = sc1; a.sc1_C = sc1_C; a.sc_stride = sc_stride;
    a.weight = weight; a.bias = bias; a.bias2 = bias2; a.residual = residual; a.relu = relu; a.out = out;
    if (use_tc == 6) {
        SYNT_CHECK(act_dtype == DT_BF16 && conv_tc2_supported(a), "conv_tc2: unsupported");
        conv_tc2(a, (cudaStream_t)stream);
    } else if (use_tc >= 2) {
#ifdef SYNT_EXPERIMENTS
        SYNT_CHECK(act_dtype == DT_BF16, "tcgen05 conv needs bf16 activations");
        conv_tc_halo(a, use_tc - 2, (cudaStream_t)stream);
#else
        throw Error(-3, "the halo-descriptor experiments (tools/exp_halo.py) need a build with SYNT_EXPERIMENTS=1");
#endif
    } else if (use_tc) {
        SYNT_CHECK(act_dtype == DT_BF16, "tcgen05 conv needs bf16 activations");
        conv_tc(a, (cudaStream_t)stream);
    } else {
        conv_simt(a, act_dtype, (cudaStream_t)stream);
    }
    SYNT_CATCH
}

// Upsample2D + conv3x3 through the fused sub-pixel path of conv_tc2: in [B,H,W,Cin] bf16, w_host the raw
// [Cout][Cin][3][3] fp32 filter (host), out [B,2H,2W,Cout] bf16, optional GroupNorm partials of `out`.
extern "C" int synt_debug_conv_up2x(const void* in, int B, int H, int W, int Cin, const float* w_host, const float* bias,
                                    void* out, int Cout, void* stats_out, int* stats_slots, void* stream) {
    SYNT_TRY
    SYNT_CHECK(in && w_host && bias && out, "bad argument");
    WeightDev wd = upload_weight(pack_upsample_phases(w_host, Cout, Cin), false, true);
    ConvArgs a;
    a.in = in; a.B = B; a.H = H; a.W = W; a.Cin = Cin; a.Ho = H; a.Wo = W; a.Cout = Cout; a.up2x = 1;
    a.weight = wd.b16->p; a.bias = bias; a.out = out; a.stats_out = (float2*)stats_out;
    SYNT_CHECK(conv_tc2_supported(a), "conv_tc2: unsupported");
    if (stats_slots) *stats_slots = conv_tc2_stats_slots(a);
    conv_tc2(a, (cudaStream_t)stream);
    SYNT_CUDA(cudaStreamSynchronize((cudaStream_t)stream));     // wd is freed on return
    SYNT_CATCH
}

// Switches the experimental CTA-pair (tcgen05 cta_group::2, cluster of two CTAs) variant of conv_tc2 on or off.
extern "C" int synt_debug_set_conv_pair(int on) { conv_tc2_set_pair(on ? 1 : 0); return 0; }

// Host-only: the phase-stacked filter of the fused Upsample2D + conv3x3 (pack_upsample_phases), for CPU tests of
// the sub-pixel decomposition.  w [Cout][Cin][3][3] -> out [4*Cout][4*Cin] (row = phase*Cout + n, col = tap*Cin + c).
extern "C" int synt_debug_pack_upsample_phases(const float* w, int Cout, int Cin, float* out) {
    SYNT_TRY
    SYNT_CHECK(w && out && Cout > 0 && Cin > 0, "bad argument");
    const std::vector<float> o = pack_upsample_phases(w, Cout, Cin);
    memcpy(out, o.data(), o.size() * sizeof(float));
    SYNT_CATCH
}

// attention core on caller-provided tensors: use_tc=0 -> qkv [B,N,3C] (q|k|v) of `act_dtype`;
// use_tc=1 -> qkv [B,N,3C] bf16 with q pre-scaled by log2(e)/sqrt(8) (see kernels.cuh)
extern "C" int synt_debug_attention(int use_tc, int act_dtype, const void* qkv, int B, int N, int C, void* out,
                                    void* stream) {
    SYNT_TRY
    SYNT_CHECK(qkv && out && B > 0, "bad argument");
    if (use_tc) {
        SYNT_CHECK(act_dtype == DT_BF16 && attention_tc_supported(N, C), "attention_tc: unsupported");
        attention_tc(qkv, B, N, C, nullptr, out, (cudaStream_t)stream);
    } else {
        attention_simt(qkv, act_dtype, B, N, C, out, (cudaStream_t)stream);
    }
    SYNT_CATCH
}

// conv_tc2 with its fused-input features: channel-concat main input (in | in1) and GroupNorm scale/shift
// (+SiLU) applied inside the kernel.  gn_ss: [B][Cin+Cin1] float2 (scale, shift); gn_mode 0/1/2.
extern "C" int synt_debug_conv_gn(const void* in, int Cin, const void* in1, int Cin1, const void* gn_ss, int gn_mode,
                                  int B, int H, int W, int K, const void* sc0, int sc0_C, const void* weight,
                                  const float* bias, const void* residual, void* out, int Cout, void* stats_out,
                                  int* stats_slots, void* stream) {
    SYNT_TRY
    ConvArgs a;
    a.in = in; a.Cin = Cin; a.in1 = in1; a.Cin1 = Cin1; a.gn_ss = (const float2*)gn_ss; a.gn_mode = gn_mode;
    a.B = B; a.H = H; a.W = W; a.KH = a.KW = K; a.pad = K / 2; a.Ho = H; a.Wo = W; a.Cout = Cout;
    a.sc0 = sc0; a.sc0_C = sc0_C; a.weight = weight; a.bias = bias; a.residual = residual; a.out = out;
    a.stats_out = (float2*)stats_out;
    SYNT_CHECK(conv_tc2_supported(a), "conv_tc2: unsupported");
    if (stats_slots) *stats_slots = conv_tc2_stats_slots(a);
    conv_tc2(a, (cudaStream_t)stream);
    SYNT_CATCH
}
