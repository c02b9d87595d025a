"""CPU check of the index arithmetic the input-gradient kernels use (synt_isic_b200/csrc/resnet_grad.cu, the dgrad weight
re-arrangement in resnet.cu::ensure_bwd), restated in numpy/torch loops and compared with torch autograd on small shapes:

* data gradient of a K-major 3x3 convolution = stride-1 convolution with Wd[n][t*Cout + co] = Wf[co][(taps-1-t)*Cin + n]
* stride-2 3x3 conv + 1x1/s2 downsample: the same over ZERO-INSERTED gradient planes, the 1x1 part as a shortcut segment
* 3x3/s2 max-pool adjoint as a gather with ATen's first-maximum rule (ties!)
* 7x7/s2 stem data gradient by parity class
* adjoint of clamp -> bilinear 128->224 (align_corners=False, antialias=True) -> normalise as a gather with the forward
  kernel's source-index arithmetic
"""
import math

import numpy as np
import torch
import torch.nn.functional as F


def kmajor(w):                               # [co][ci][kh][kw] -> [co][t*ci + ci]   (ConvArgs weight layout)
    co, ci, kh, kw = w.shape
    return w.permute(0, 2, 3, 1).reshape(co, kh * kw * ci)


def from_kmajor(m, cin, k):
    return m.reshape(m.shape[0], k, k, cin).permute(0, 3, 1, 2).contiguous()


def dgrad_w(Wf, src_off, cin_f, cout_f, taps):
    """dgrad_weight_kernel"""
    Wd = torch.zeros(cin_f, taps * cout_f)
    for n in range(cin_f):
        for t in range(taps):
            for co in range(cout_f):
                Wd[n, t * cout_f + co] = Wf[co, src_off + (taps - 1 - t) * cin_f + n]
    return Wd


def zero_insert(g):
    B, C, H, W = g.shape
    o = torch.zeros(B, C, 2 * H, 2 * W)
    o[:, :, ::2, ::2] = g
    return o


def test_dgrad_weight_matrices():
    torch.manual_seed(0)
    ci, c = 4, 6
    x = torch.randn(2, ci, 8, 8, requires_grad=True)
    w = torch.randn(c, ci, 3, 3)
    y = F.conv2d(x, w, padding=1)
    gy = torch.randn_like(y)
    y.backward(gy)
    dx = F.conv2d(gy, from_kmajor(dgrad_w(kmajor(w), 0, ci, c, 9), c, 3), padding=1)
    assert (dx - x.grad).abs().max().item() < 1e-4
    # stride-2 block: conv1 3x3/s2 + downsample 1x1/s2 read the same input
    x = torch.randn(2, ci, 8, 8, requires_grad=True)
    w1, wds = torch.randn(c, ci, 3, 3), torch.randn(c, ci, 1, 1)
    y1, y2 = F.conv2d(x, w1, stride=2, padding=1), F.conv2d(x, wds, stride=2)
    g1, g2 = torch.randn_like(y1), torch.randn_like(y2)
    ((y1 * g1).sum() + (y2 * g2).sum()).backward()
    Wf2 = torch.cat([torch.randn(c, 9 * c), kmajor(wds)], 1)                    # [c][9c + ci] like c2[l][0]
    Wd1 = torch.cat([dgrad_w(kmajor(w1), 0, ci, c, 9), dgrad_w(Wf2, 9 * c, ci, c, 1)], 1)
    dx = (F.conv2d(zero_insert(g1), from_kmajor(Wd1[:, :9 * c], c, 3), padding=1)
          + F.conv2d(zero_insert(g2), from_kmajor(Wd1[:, 9 * c:], c, 1)))
    assert (dx - x.grad).abs().max().item() < 1e-4


def test_maxpool_gather_with_ties():
    torch.manual_seed(1)
    H = W = 10
    Ho = Wo = 5
    a = torch.randint(0, 4, (1, 2, H, W)).float().requires_grad_(True)          # many equal values
    p = F.max_pool2d(a, 3, 2, 1)
    gp = torch.randn_like(p)
    p.backward(gp)
    A, da = a.detach(), torch.zeros(1, 2, H, W)
    for ch in range(2):
        for iy in range(H):
            for ix in range(W):
                s, acc = A[0, ch, iy, ix], 0.0
                for oy in range(iy >> 1, ((iy + 1) >> 1) + 1):
                    for ox in range(ix >> 1, ((ix + 1) >> 1) + 1):
                        if oy >= Ho or ox >= Wo:
                            continue
                        my, mx, win = iy - (2 * oy - 1), ix - (2 * ox - 1), True
                        for dy in range(3):
                            for dx in range(3):
                                yy, xx = 2 * oy - 1 + dy, 2 * ox - 1 + dx
                                if yy < 0 or yy >= H or xx < 0 or xx >= W or (dy == my and dx == mx):
                                    continue
                                o = A[0, ch, yy, xx]
                                before = dy < my or (dy == my and dx < mx)
                                win = win and bool((o < s) if before else (o <= s))
                        if win:
                            acc += gp[0, ch, oy, ox].item()
                da[0, ch, iy, ix] = acc
    assert (da - a.grad).abs().max().item() < 1e-6


def test_stem_dgrad_by_parity_class():
    torch.manual_seed(2)
    x = torch.randn(1, 3, 16, 16, requires_grad=True)
    w = torch.randn(5, 3, 7, 7)
    y = F.conv2d(x, w, stride=2, padding=3)
    g = torch.randn_like(y)
    y.backward(g)
    wk, Ho = kmajor(w), y.shape[2]
    dx = torch.zeros(1, 3, 16, 16)
    for iy in range(16):
        for ix in range(16):
            ky0, kx0 = ((iy & 1) + 1) & 1, ((ix & 1) + 1) & 1
            for a in range(3 if ky0 else 4):
                ny = iy + 3 - (ky0 + 2 * a)
                if ny < 0 or (ny >> 1) >= Ho:
                    continue
                for e in range(3 if kx0 else 4):
                    nx = ix + 3 - (kx0 + 2 * e)
                    if nx < 0 or (nx >> 1) >= Ho:
                        continue
                    ky, kx = ky0 + 2 * a, kx0 + 2 * e
                    for c in range(3):
                        dx[0, c, iy, ix] += (g[0, :, ny >> 1, nx >> 1] * wk[:, (ky * 7 + kx) * 3 + c]).sum()
    assert (dx - x.grad).abs().max().item() < 1e-4


def test_preprocess_adjoint_gather():
    torch.manual_seed(3)
    f32 = np.float32
    x = (torch.randn(1, 3, 128, 128) * 0.7).requires_grad_(True)
    z = F.interpolate(torch.clamp((x + 1) / 2, 0, 1), size=(224, 224), mode="bilinear", align_corners=False, antialias=True)
    std = torch.tensor([0.229, 0.224, 0.225]).view(1, 3, 1, 1)
    z = (z - torch.tensor([0.485, 0.456, 0.406]).view(1, 3, 1, 1)) / std
    gz = torch.randn_like(z)
    z.backward(gz)
    sc = f32(128) / f32(224)
    Wm = np.zeros((128, 224), dtype=np.float32)
    for pos in range(128):                                                       # preprocess_bwd_kernel::weights
        lo = max(int(math.floor(f32((f32(pos) - f32(0.5)) / sc - f32(0.5)))) - 1, 0)
        for k in range(8):
            o = lo + k
            if o >= 224:
                continue
            f = max(f32((f32(o) + f32(0.5)) * sc - f32(0.5)), f32(0))
            p0 = int(f)
            p1 = p0 + (1 if p0 < 127 else 0)
            lam = f32(f - f32(p0))
            Wm[pos, o] += (f32(1) - lam if pos == p0 else 0) + (lam if pos == p1 else 0)
    assert np.allclose(Wm.sum(0), 1.0, atol=1e-6)                               # every output pixel fully distributed
    Wt = torch.from_numpy(Wm)
    dx = torch.einsum("yo,bcop,xp->bcyx", Wt, gz / std, Wt) * 0.5
    v = (x.detach() + 1) / 2
    dx = torch.where((v >= 0) & (v <= 1), dx, torch.zeros_like(dx))
    assert (dx - x.grad).abs().max().item() < 1e-4
