"""Kernel-level parity of the two GEMM carriers (tcgen05 implicit GEMM, fp32-FMA implicit GEMM)
against torch.nn.functional.conv2d on the same (bf16-rounded) operands, through the C ABI."""
import ctypes as C

import pytest
import torch
import torch.nn.functional as F

from synt_isic_b200 import _lib

pytestmark = pytest.mark.gpu


def _pack(w, wsc=None):
    """[Cout,Cin,k,k] (+ [Cout,Csc]) -> K-major [Cout, k*k*Cin + Csc] (tap-major, channel-minor)."""
    cout = w.shape[0]
    p = w.permute(0, 2, 3, 1).reshape(cout, -1)
    if wsc is not None:
        p = torch.cat([p, wsc.reshape(cout, -1)], dim=1)
    return p.contiguous()


def run_conv(dev, mode, B, H, W, Cin, Cout, K, stride, pad, sc=(0, 0), sc_stride=1, residual=False, relu=False,
             bias2=False, seed=0):
    use_tc = mode != 0
    g = torch.Generator().manual_seed(seed)
    dt = torch.bfloat16 if use_tc else torch.float32
    x = torch.randn(B, Cin, H, W, generator=g)
    w = torch.randn(Cout, Cin, K, K, generator=g) / (Cin * K * K) ** 0.5
    b = torch.randn(Cout, generator=g)
    Ho, Wo = (H + 2 * pad - K) // stride + 1, (W + 2 * pad - K) // stride + 1
    xs = [torch.randn(B, c, Ho * sc_stride, Wo * sc_stride, generator=g) if c else None for c in sc]
    ws = [torch.randn(Cout, c, generator=g) / c ** 0.5 if c else None for c in sc]
    res = torch.randn(B, Cout, Ho, Wo, generator=g) if residual else None
    b2 = torch.randn(Cout, generator=g) if bias2 else None
    q = (lambda t: t.to(dt).float()) if use_tc else (lambda t: t)
    # reference on the same rounded operands, fp32 math on the GPU (TF32 off)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    ref = F.conv2d(q(x).to(dev), q(w).to(dev), b.to(dev), stride=stride, padding=pad)
    for xi, wi in zip(xs, ws):
        if xi is not None:
            ref = ref + F.conv2d(q(xi).to(dev)[:, :, ::sc_stride, ::sc_stride], q(wi).to(dev)[:, :, None, None])
    if b2 is not None:
        ref = ref + b2.to(dev)[None, :, None, None]
    if res is not None:
        ref = ref + q(res).to(dev)
    if relu:
        ref = ref.relu()
    nhwc = lambda t: t.permute(0, 2, 3, 1).contiguous().to(dt).to(dev) if t is not None else None
    xin, xs0, xs1, rs = nhwc(x), nhwc(xs[0]), nhwc(xs[1]), nhwc(res)
    wsc = torch.cat([v for v in ws if v is not None], dim=1) if any(v is not None for v in ws) else None
    wp = _pack(w, wsc).to(dt).to(dev)
    bd = b.to(dev)
    b2d = b2.to(dev) if b2 is not None else None
    out = torch.empty(B, Ho, Wo, Cout, dtype=dt, device=dev)
    ptr = lambda t: t.data_ptr() if t is not None else None
    _lib.check(_lib.lib().synt_debug_conv(int(mode), 1 if use_tc else 0, ptr(xin), B, H, W, Cin, K, stride, pad,
                                          ptr(xs0), sc[0], ptr(xs1), sc[1], sc_stride, ptr(wp), ptr(bd), ptr(b2d),
                                          ptr(rs), 1 if relu else 0, ptr(out), Cout, _lib.current_stream_ptr()),
               "debug_conv")
    torch.cuda.synchronize()
    got = out.float().permute(0, 3, 1, 2)
    return ((got - ref).norm() / ref.norm()).item(), (got - ref).abs().max().item()


CASES = [
    # B, H, W, Cin, Cout, K, stride, pad, sc, sc_stride, residual, relu, bias2
    (2, 32, 32, 64, 64, 3, 1, 1, (0, 0), 1, False, False, False),
    (1, 128, 128, 64, 64, 3, 1, 1, (0, 0), 1, True, False, True),
    (2, 16, 16, 256, 256, 3, 1, 1, (0, 0), 1, True, False, True),
    (2, 32, 32, 128, 256, 3, 1, 1, (128, 0), 1, False, False, False),      # fused 1x1 shortcut
    (1, 32, 32, 256, 256, 3, 1, 1, (256, 128), 1, False, False, True),     # concat shortcut (384 -> 256)
    (2, 64, 64, 128, 128, 3, 2, 1, (0, 0), 1, False, False, False),        # Downsample2D
    (2, 32, 32, 256, 768, 1, 1, 0, (0, 0), 1, False, False, False),        # qkv projection
    (2, 32, 32, 256, 256, 1, 1, 0, (0, 0), 1, True, False, False),         # out projection + residual
    (3, 56, 56, 64, 64, 3, 1, 1, (0, 0), 1, True, True, False),            # ResNet18 layer1 (ragged 56)
    (3, 56, 56, 64, 128, 3, 2, 1, (0, 0), 1, False, True, False),          # ResNet18 layer2.0.conv1
    (3, 28, 28, 128, 128, 3, 1, 1, (64, 0), 2, False, True, False),        # layer2.0.conv2 + 1x1 s2 downsample
    (3, 56, 56, 64, 128, 1, 2, 0, (0, 0), 1, False, False, False),         # layer2.0 downsample as its own 1x1 / stride-2 conv
    (3, 28, 28, 128, 128, 3, 1, 1, (0, 0), 1, True, True, False),          # layer2.0.conv2 with the downsample output as residual
    (3, 14, 14, 256, 256, 3, 1, 1, (0, 0), 1, True, True, False),          # 14x14
    (3, 7, 7, 512, 512, 3, 1, 1, (0, 0), 1, True, True, False),            # 7x7, two images per tile, odd B
    (16, 32, 32, 256, 256, 3, 1, 1, (0, 0), 1, True, False, True),         # 128 pixel tiles x 2 -> N tile 128
    (40, 32, 32, 128, 256, 3, 1, 1, (128, 0), 1, False, False, False),     # 320 pixel tiles -> N tile 256
    (10, 64, 64, 128, 128, 3, 1, 1, (0, 0), 1, True, False, False),        # N tile 128, 3 smem stages
    (20, 32, 32, 256, 1280, 1, 1, 0, (0, 0), 1, False, False, False),      # zero-interleaved qkv projection, N tile 256
]


@pytest.mark.parametrize("case", CASES, ids=lambda c: "x".join(map(str, c[:8])))
def test_conv_simt_fp32(cuda_dev, case):
    B, H, W, Cin, Cout, K, s, p, sc, scs, res, relu, b2 = case
    rel, mx = run_conv(cuda_dev, False, B, H, W, Cin, Cout, K, s, p, sc, scs, res, relu, b2)
    assert rel < 5e-6, (rel, mx)        # fp32 tolerance: accumulation-order differences only


@pytest.mark.parametrize("case", CASES, ids=lambda c: "x".join(map(str, c[:8])))
def test_conv_tcgen05_bf16(cuda_dev, case):
    B, H, W, Cin, Cout, K, s, p, sc, scs, res, relu, b2 = case
    rel, mx = run_conv(cuda_dev, True, B, H, W, Cin, Cout, K, s, p, sc, scs, res, relu, b2)
    assert rel < 4e-3, (rel, mx)        # operands identical; only the bf16 rounding of the OUTPUT differs


def test_conv_simt_stem_7x7(cuda_dev):
    rel, mx = run_conv(cuda_dev, False, 2, 224, 224, 3, 64, 7, 2, 3, relu=True)
    assert rel < 2e-6, (rel, mx)


V2_CASES = [c for c in CASES if c[5] in (1, 3) and c[6] == 1 and c[1] % 16 == 0 and (c[1] % 32 == 0 or c[1] == 16) and c[9] == 1] + [
    (3, 16, 16, 512, 256, 3, 1, 1, (256, 256), 1, False, False, True),     # up_blocks.0: concat shortcut, odd B, 2 images/super-tile
    (5, 64, 64, 192, 128, 3, 1, 1, (128, 64), 1, False, False, True),      # up_blocks.2.resnets.2 conv2-like
    (2, 128, 128, 128, 64, 3, 1, 1, (0, 0), 1, True, False, False),        # persistent: 256 super-tiles on 148 CTAs
    (3, 56, 56, 64, 64, 3, 1, 1, (0, 0), 1, True, True, False),            # ResNet18 layer1: ragged 56x56 plane (tile grid rounded up)
    (3, 28, 28, 128, 128, 3, 1, 1, (0, 0), 1, True, True, False),          # ResNet18 layer2: ragged 28x28 plane
    (2, 40, 24, 64, 128, 3, 1, 1, (64, 0), 1, False, False, True),         # ragged in both directions + shortcut segment
    (3, 14, 14, 256, 256, 3, 1, 1, (0, 0), 1, True, True, False),          # ResNet18 layer3: 14x14 in the two-images-per-super-tile mode
]


@pytest.mark.parametrize("case", V2_CASES, ids=lambda c: "x".join(map(str, c[:8])))
def test_conv_tcgen05_persistent_halo(cuda_dev, case):
    """conv_tc2: persistent halo-tile kernel (3x3 stride 1)."""
    B, H, W, Cin, Cout, K, s, p, sc, scs, res, relu, b2 = case
    rel, mx = run_conv(cuda_dev, 6, B, H, W, Cin, Cout, K, s, p, sc, scs, res, relu, b2)
    assert rel < 4e-3, (rel, mx)


GN_CASES = [
    # B, H, W, Cin, Cin1, Cout, K, gn_mode, sc0_C, residual
    (2, 32, 32, 64, 0, 64, 3, 2, 0, True),          # resnet conv2 (identity residual), resident weights
    (3, 16, 16, 256, 256, 256, 3, 2, 0, False),     # up_blocks.0 conv1: concat input, 2 images / super-tile, odd B
    (2, 64, 64, 128, 64, 128, 3, 2, 0, False),      # up_blocks.2.resnets.2 conv1: 192-channel concat
    (2, 32, 32, 256, 0, 256, 3, 2, 128, False),     # conv2 + raw (untransformed) 1x1 shortcut segment
    (2, 32, 32, 256, 0, 1280, 1, 1, 0, False),      # attention qkv projection: GroupNorm without SiLU
    (1, 128, 128, 64, 0, 64, 3, 2, 0, True),        # 128x128: padding rows/cols must stay zero after the transform
    (2, 32, 32, 64, 0, 64, 3, 2, 128, False),       # up_blocks.3 conv2 + 1x1 shortcut over 128 channels (Cout = 64, K = 704)
    (3, 64, 64, 64, 0, 64, 3, 2, 192, False),       # ... over 192 channels (K = 768), more work items than slots
]


@pytest.mark.parametrize("case", GN_CASES, ids=lambda c: "x".join(map(str, c)))
def test_conv_tcgen05_fused_groupnorm_input(cuda_dev, case):
    """conv_tc2 with GroupNorm(+SiLU) applied to the input inside the kernel and fused output statistics."""
    B, H, W, Cin, Cin1, Cout, K, mode, scC, res = case
    dev = cuda_dev
    g = torch.Generator().manual_seed(sum(case))
    Ct = Cin + Cin1
    x = torch.randn(B, Ct, H, W, generator=g) * 1.5 + 0.3
    sc = 0.5 + torch.rand(B, Ct, generator=g)
    sh = torch.randn(B, Ct, generator=g) * 0.5
    w = torch.randn(Cout, Ct, K, K, generator=g) / (Ct * K * K) ** 0.5
    b = torch.randn(Cout, generator=g)
    xs = torch.randn(B, scC, H, W, generator=g) if scC else None
    ws = torch.randn(Cout, scC, generator=g) / scC ** 0.5 if scC else None
    r = torch.randn(B, Cout, H, W, generator=g) if res else None
    q = lambda t: t.to(torch.bfloat16).float()
    xq = q(x).to(dev)
    y = xq * sc.to(dev)[:, :, None, None] + sh.to(dev)[:, :, None, None]
    if mode == 2:
        y = F.silu(y)
    y = q(y)                                                       # the kernel rounds the transformed tile to bf16
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    ref = F.conv2d(y, q(w).to(dev), b.to(dev), padding=K // 2)
    if scC:
        ref = ref + F.conv2d(q(xs).to(dev), q(ws).to(dev)[:, :, None, None])
    if res:
        ref = ref + q(r).to(dev)
    nhwc = lambda t: t.permute(0, 2, 3, 1).contiguous().to(torch.bfloat16).to(dev) if t is not None else None
    x0, x1 = nhwc(x[:, :Cin]), (nhwc(x[:, Cin:]) if Cin1 else None)
    wp = _pack(w, ws).to(torch.bfloat16).to(dev)
    ss = torch.stack([sc, sh], dim=2).contiguous().to(dev)        # [B][Ct] float2
    out = torch.empty(B, H, W, Cout, dtype=torch.bfloat16, device=dev)
    stats = torch.zeros(B * 256 * Cout * 2, dtype=torch.float32, device=dev)
    slots = C.c_int()
    ptr = lambda t: t.data_ptr() if t is not None else None
    xs_d, r_d, b_d = nhwc(xs), nhwc(r), b.to(dev)                 # keep the device tensors alive across the call
    _lib.check(_lib.lib().synt_debug_conv_gn(ptr(x0), Cin, ptr(x1), Cin1, ptr(ss), mode, B, H, W, K, ptr(xs_d), scC, ptr(wp),
                                             ptr(b_d), ptr(r_d), ptr(out), Cout, ptr(stats), C.byref(slots),
                                             _lib.current_stream_ptr()), "debug_conv_gn")
    torch.cuda.synchronize()
    got = out.float().permute(0, 3, 1, 2)
    rel = ((got - ref).norm() / ref.norm()).item()
    assert rel < 6e-3, rel                                         # tanh.approx SiLU + bf16 re-rounding of the input
    # fused output statistics == per-channel sums of the stored bf16 output
    st = stats[: B * slots.value * Cout * 2].view(B, slots.value, Cout, 2).sum(1)
    want_s = got.sum(dim=(2, 3)); want_q = (got * got).sum(dim=(2, 3))
    assert torch.allclose(st[..., 0], want_s, rtol=2e-3, atol=2e-2)
    assert torch.allclose(st[..., 1], want_q, rtol=2e-3, atol=2e-2)


K1_CASES = [
    # B, H, W, Cin, Cout, gn_mode      (no residual, no statistics: the no-halo four-slot variant of conv_tc2)
    (2, 32, 32, 256, 768, 1),          # qkv projection with the attention block's GroupNorm affine fused
    (3, 16, 16, 256, 768, 1),          # 16x16 site: two images per super-tile, odd batch
    (2, 32, 32, 256, 768, 0),          # raw input (TMA only)
    (70, 32, 32, 256, 768, 1),         # more work items than SMs x ring depth
]


@pytest.mark.parametrize("case", K1_CASES, ids=lambda c: "x".join(map(str, c)))
def test_conv_tcgen05_1x1_four_slot_variant(cuda_dev, case):
    """The qkv 1x1 projection path of the sampling step (diffusers Attention.to_q/k/v behind group_norm, reached from
    core/generator/image_generator.py:400): 1x1 conv with the GroupNorm affine applied to the input inside the kernel."""
    B, H, W, Cin, Cout, mode = case
    dev = cuda_dev
    g = torch.Generator().manual_seed(sum(case))
    x = torch.randn(B, Cin, H, W, generator=g) * 1.5 + 0.3
    sc = 0.5 + torch.rand(B, Cin, generator=g)
    sh = torch.randn(B, Cin, generator=g) * 0.5
    w = torch.randn(Cout, Cin, 1, 1, generator=g) / Cin ** 0.5
    b = torch.randn(Cout, generator=g)
    q = lambda t: t.to(torch.bfloat16).float()
    y = q(x).to(dev)
    if mode:
        y = q(y * sc.to(dev)[:, :, None, None] + sh.to(dev)[:, :, None, None])
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    ref = F.conv2d(y, q(w).to(dev), b.to(dev))
    x0 = x.permute(0, 2, 3, 1).contiguous().to(torch.bfloat16).to(dev)
    wp = _pack(w, None).to(torch.bfloat16).to(dev)
    ss = torch.stack([sc, sh], dim=2).contiguous().to(dev)
    out = torch.empty(B, H, W, Cout, dtype=torch.bfloat16, device=dev)
    slots = C.c_int()
    b_d = b.to(dev)
    _lib.check(_lib.lib().synt_debug_conv_gn(x0.data_ptr(), Cin, None, 0, ss.data_ptr() if mode else None, mode, B, H, W, 1, None, 0,
                                             wp.data_ptr(), b_d.data_ptr(), None, out.data_ptr(), Cout, None, C.byref(slots),
                                             _lib.current_stream_ptr()), "debug_conv_gn")
    torch.cuda.synchronize()
    got = out.float().permute(0, 3, 1, 2)
    rel = ((got - ref).norm() / ref.norm()).item()
    assert rel < 4e-3, rel


UP_CASES = [
    # B, H, W (low-res), Cin, Cout
    (3, 16, 16, 256, 256),      # up_blocks.0 upsampler: 2 images per super-tile, odd B
    (2, 32, 32, 256, 256),      # up_blocks.1 upsampler
    (2, 64, 64, 128, 128),      # up_blocks.2 upsampler
    (1, 32, 32, 64, 64),        # Cout = 64 (N tile 64)
]


@pytest.mark.parametrize("case", UP_CASES, ids=lambda c: "x".join(map(str, c)))
def test_conv_tcgen05_fused_upsample(cuda_dev, case):
    """Upsample2D (nearest 2x) + conv3x3 as four sub-pixel 2x2 convolutions inside conv_tc2
    (diffusers Upsample2D reached from core/generator/image_generator.py:400) vs F.interpolate + F.conv2d."""
    B, H, W, Cin, Cout = case
    dev = cuda_dev
    g = torch.Generator().manual_seed(sum(case))
    x = torch.randn(B, Cin, H, W, generator=g)
    w = torch.randn(Cout, Cin, 3, 3, generator=g) / (Cin * 9) ** 0.5
    b = torch.randn(Cout, generator=g)
    q = lambda t: t.to(torch.bfloat16).float()
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    # the reference order of operations, fp32 weights (the fused path sums taps in fp32 and rounds once)
    ref = F.conv2d(F.interpolate(q(x).to(dev), scale_factor=2.0, mode="nearest"), w.to(dev), b.to(dev), padding=1)
    xin = x.permute(0, 2, 3, 1).contiguous().to(torch.bfloat16).to(dev)
    out = torch.empty(B, 2 * H, 2 * W, Cout, dtype=torch.bfloat16, device=dev)
    stats = torch.zeros(B * 1024 * Cout * 2, dtype=torch.float32, device=dev)
    slots = C.c_int()
    wh = w.contiguous()
    bd = b.to(dev)
    _lib.check(_lib.lib().synt_debug_conv_up2x(xin.data_ptr(), B, H, W, Cin, wh.data_ptr(), bd.data_ptr(), out.data_ptr(),
                                               Cout, stats.data_ptr(), C.byref(slots), _lib.current_stream_ptr()),
               "debug_conv_up2x")
    torch.cuda.synchronize()
    got = out.float().permute(0, 3, 1, 2)
    rel = ((got - ref).norm() / ref.norm()).item()
    assert rel < 5e-3, rel                                         # bf16 rounding of the (pre-summed) weights and of the output
    st = stats[: B * slots.value * Cout * 2].view(B, slots.value, Cout, 2).sum(1)
    want_s = got.sum(dim=(2, 3)); want_q = (got * got).sum(dim=(2, 3))
    assert torch.allclose(st[..., 0], want_s, rtol=2e-3, atol=5e-2)
    assert torch.allclose(st[..., 1], want_q, rtol=2e-3, atol=5e-2)


PAIR_CASES = [c for c in V2_CASES if c[4] % 128 == 0 and not (c[4] % 256 == 0 and c[1] >= 32 and c[5] == 3 and c[3] >= 128)]


@pytest.mark.parametrize("case", PAIR_CASES, ids=lambda c: "x".join(map(str, c[:8])))
def test_conv_tcgen05_cta_pair(cuda_dev, case):
    """Experimental CTA-pair variant of conv_tc2 (cluster of two CTAs, tcgen05.mma.cta_group::2 with M = 256, each CTA
    loads half of every weight tile, multicast commits, relayed barriers): same results as the single-CTA kernel."""
    B, H, W, Cin, Cout, K, s, p, sc, scs, res, relu, b2 = case
    _lib.check(_lib.lib().synt_debug_set_conv_pair(1), "set_conv_pair")
    try:
        rel, mx = run_conv(cuda_dev, 6, B, H, W, Cin, Cout, K, s, p, sc, scs, res, relu, b2)
    finally:
        _lib.check(_lib.lib().synt_debug_set_conv_pair(0), "set_conv_pair")
    assert rel < 4e-3, (rel, mx)
