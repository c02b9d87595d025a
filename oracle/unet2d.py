"""TEST INFRASTRUCTURE ONLY (see oracle/__init__.py) -- **parity unpinned** vs real diffusers.

Plain-PyTorch fp32 restatement of ``diffusers.UNet2DModel`` for exactly the constructor
arguments the reference passes (``core/generator/model_manager.py:175-194``, repeated at
``core/generator/image_generator.py:266-288``, ``diffusion/diffusion_generator.py:75-93``,
``xai/XAI.py:313-339``); every other argument is the diffusers default.

Module/parameter names reproduce the diffusers ``state_dict`` key scheme
so that ``load_state_dict(strict=True)`` (``xai/XAI.py:605``) round-trips, and the total
parameter count must equal 25,304,963 (shipped checkpoint size pin,
``core/cache/metadata/cache_metadata.json:7``).

Every ambiguous diffusers choice sits behind a named constant so a later check against
the real package can flip it.
"""
from __future__ import annotations

import math
from types import SimpleNamespace

import torch
import torch.nn as nn
import torch.nn.functional as F

# ---- named constants (diffusers defaults for the reference's constructor call) -------
SAMPLE_SIZE = 128                       # model_manager.py:176
IN_CHANNELS = 3                         # model_manager.py:177
OUT_CHANNELS = 3                        # model_manager.py:178
LAYERS_PER_BLOCK = 2                    # model_manager.py:179
BLOCK_OUT_CHANNELS = (64, 128, 256, 256)  # model_manager.py:180
DOWN_HAS_ATTN = (False, False, True, False)   # Down, Down, AttnDown, Down  (:181-186)
UP_HAS_ATTN = (False, True, False, False)     # Up, AttnUp, Up, Up          (:187-192)
ATTENTION_HEAD_DIM = 8                  # diffusers default
NORM_NUM_GROUPS = 32                    # diffusers default
NORM_EPS = 1e-5                         # diffusers default
FLIP_SIN_TO_COS = True                  # diffusers default -> [cos, sin]
FREQ_SHIFT = 0                          # diffusers default
TIME_EMBED_DIM = BLOCK_OUT_CHANNELS[0] * 4    # 256
RESNET_TIME_SCALE_SHIFT = "default"     # additive temb (checkpoint-size pin)
OUTPUT_SCALE_FACTOR = 1.0
EXPECTED_PARAM_COUNT = 25_304_963
# "explicit": softmax(q k^T / sqrt(d)) v spelled out (parity runs).  "sdpa": F.scaled_dot_product_attention, what diffusers'
# AttnProcessor2_0 calls -- used by bench.py's same-box GPU eager baseline so that the baseline gets torch's fused kernels.
ATTENTION_IMPL = "explicit"


def timestep_embedding(timesteps: torch.Tensor, dim: int = BLOCK_OUT_CHANNELS[0]) -> torch.Tensor:
    """diffusers ``get_timestep_embedding`` (embeddings.py) with flip_sin_to_cos=True,
    downscale_freq_shift=0, scale=1, max_period=10000."""
    half = dim // 2
    exponent = -math.log(10000.0) * torch.arange(half, dtype=torch.float32, device=timesteps.device)
    exponent = exponent / (half - FREQ_SHIFT)
    emb = timesteps[:, None].float() * torch.exp(exponent)[None, :]
    emb = torch.cat([torch.sin(emb), torch.cos(emb)], dim=-1)
    if FLIP_SIN_TO_COS:
        emb = torch.cat([emb[:, half:], emb[:, :half]], dim=-1)
    return emb


class TimestepEmbedding(nn.Module):
    def __init__(self, in_dim: int, dim: int):
        super().__init__()
        self.linear_1 = nn.Linear(in_dim, dim)
        self.linear_2 = nn.Linear(dim, dim)

    def forward(self, x):
        return self.linear_2(F.silu(self.linear_1(x)))


class ResnetBlock2D(nn.Module):
    """diffusers resnet.py::ResnetBlock2D, time_embedding_norm="default", dropout 0."""

    def __init__(self, cin: int, cout: int):
        super().__init__()
        self.norm1 = nn.GroupNorm(NORM_NUM_GROUPS, cin, eps=NORM_EPS, affine=True)
        self.conv1 = nn.Conv2d(cin, cout, 3, padding=1)
        self.time_emb_proj = nn.Linear(TIME_EMBED_DIM, cout)
        self.norm2 = nn.GroupNorm(NORM_NUM_GROUPS, cout, eps=NORM_EPS, affine=True)
        self.conv2 = nn.Conv2d(cout, cout, 3, padding=1)
        self.conv_shortcut = nn.Conv2d(cin, cout, 1) if cin != cout else None

    def forward(self, x, temb):
        h = self.conv1(F.silu(self.norm1(x)))
        h = h + self.time_emb_proj(F.silu(temb))[:, :, None, None]
        h = self.conv2(F.silu(self.norm2(h)))
        if self.conv_shortcut is not None:
            x = self.conv_shortcut(x)
        return (x + h) / OUTPUT_SCALE_FACTOR


class Attention(nn.Module):
    """diffusers attention_processor.py::Attention (+AttnProcessor2_0) as UNet2D blocks
    build it: heads=C/8, group_norm(32), bias, residual_connection, rescale 1."""

    def __init__(self, channels: int):
        super().__init__()
        self.heads = channels // ATTENTION_HEAD_DIM
        self.group_norm = nn.GroupNorm(NORM_NUM_GROUPS, channels, eps=NORM_EPS, affine=True)
        self.to_q = nn.Linear(channels, channels)
        self.to_k = nn.Linear(channels, channels)
        self.to_v = nn.Linear(channels, channels)
        self.to_out = nn.ModuleList([nn.Linear(channels, channels), nn.Dropout(0.0)])

    def forward(self, x):
        b, c, hh, ww = x.shape
        residual = x
        h = self.group_norm(x.view(b, c, hh * ww)).transpose(1, 2)       # [B, HW, C]
        q, k, v = self.to_q(h), self.to_k(h), self.to_v(h)
        d = c // self.heads
        q = q.view(b, -1, self.heads, d).transpose(1, 2)
        k = k.view(b, -1, self.heads, d).transpose(1, 2)
        v = v.view(b, -1, self.heads, d).transpose(1, 2)
        if ATTENTION_IMPL == "sdpa":
            o = F.scaled_dot_product_attention(q, k, v)
        else:
            # explicit fp32 softmax(q k^T / sqrt(d)) v  (== F.scaled_dot_product_attention)
            s = torch.matmul(q, k.transpose(-1, -2)) * (1.0 / math.sqrt(d))
            p = torch.softmax(s, dim=-1)
            o = torch.matmul(p, v)
        o = o.transpose(1, 2).reshape(b, -1, c)
        o = self.to_out[0](o)
        o = o.transpose(-1, -2).reshape(b, c, hh, ww)
        return (o + residual) / 1.0


class Downsample2D(nn.Module):
    def __init__(self, ch: int):
        super().__init__()
        self.conv = nn.Conv2d(ch, ch, 3, stride=2, padding=1)

    def forward(self, x):
        return self.conv(x)


class Upsample2D(nn.Module):
    def __init__(self, ch: int):
        super().__init__()
        self.conv = nn.Conv2d(ch, ch, 3, padding=1)

    def forward(self, x):
        return self.conv(F.interpolate(x, scale_factor=2.0, mode="nearest"))


class DownBlock(nn.Module):
    def __init__(self, cin, cout, attn: bool, add_downsample: bool):
        super().__init__()
        self.resnets = nn.ModuleList(
            [ResnetBlock2D(cin if i == 0 else cout, cout) for i in range(LAYERS_PER_BLOCK)])
        if attn:
            self.attentions = nn.ModuleList([Attention(cout) for _ in range(LAYERS_PER_BLOCK)])
        self.has_attn = attn
        if add_downsample:
            self.downsamplers = nn.ModuleList([Downsample2D(cout)])
        self.has_down = add_downsample

    def forward(self, h, temb):
        outs = []
        for i, res in enumerate(self.resnets):
            h = res(h, temb)
            if self.has_attn:
                h = self.attentions[i](h)
            outs.append(h)
        if self.has_down:
            h = self.downsamplers[0](h)
            outs.append(h)
        return h, outs


class MidBlock(nn.Module):
    def __init__(self, ch):
        super().__init__()
        self.attentions = nn.ModuleList([Attention(ch)])
        self.resnets = nn.ModuleList([ResnetBlock2D(ch, ch), ResnetBlock2D(ch, ch)])

    def forward(self, h, temb):
        h = self.resnets[0](h, temb)
        h = self.attentions[0](h)
        return self.resnets[1](h, temb)


class UpBlock(nn.Module):
    def __init__(self, cin, prev, cout, attn: bool, add_upsample: bool):
        super().__init__()
        n = LAYERS_PER_BLOCK + 1
        res = []
        for i in range(n):
            skip_ch = cin if i == n - 1 else cout
            in_ch = prev if i == 0 else cout
            res.append(ResnetBlock2D(in_ch + skip_ch, cout))
        self.resnets = nn.ModuleList(res)
        if attn:
            self.attentions = nn.ModuleList([Attention(cout) for _ in range(n)])
        self.has_attn = attn
        if add_upsample:
            self.upsamplers = nn.ModuleList([Upsample2D(cout)])
        self.has_up = add_upsample

    def forward(self, h, skips, temb):
        for i, res in enumerate(self.resnets):
            h = torch.cat([h, skips.pop()], dim=1)          # h first, then the skip
            h = res(h, temb)
            if self.has_attn:
                h = self.attentions[i](h)
        if self.has_up:
            h = self.upsamplers[0](h)
        return h


class UNet2DOracle(nn.Module):
    """``model(sample, timestep).sample`` -- call protocol of image_generator.py:400."""

    def __init__(self):
        super().__init__()
        ch = BLOCK_OUT_CHANNELS
        self.conv_in = nn.Conv2d(IN_CHANNELS, ch[0], 3, padding=1)
        self.time_embedding = TimestepEmbedding(ch[0], TIME_EMBED_DIM)
        downs, out = [], ch[0]
        for i, c in enumerate(ch):
            inp, out = out, c
            downs.append(DownBlock(inp, out, DOWN_HAS_ATTN[i], add_downsample=i != len(ch) - 1))
        self.down_blocks = nn.ModuleList(downs)
        self.mid_block = MidBlock(ch[-1])
        rev = tuple(reversed(ch))
        ups, out = [], rev[0]
        for i in range(len(rev)):
            prev, out = out, rev[i]
            inp = rev[min(i + 1, len(rev) - 1)]
            ups.append(UpBlock(inp, prev, out, UP_HAS_ATTN[i], add_upsample=i != len(rev) - 1))
        self.up_blocks = nn.ModuleList(ups)
        self.conv_norm_out = nn.GroupNorm(NORM_NUM_GROUPS, ch[0], eps=NORM_EPS)
        self.conv_out = nn.Conv2d(ch[0], OUT_CHANNELS, 3, padding=1)

    @property
    def device(self):
        return next(self.parameters()).device

    def time_embed(self, timestep, batch: int) -> torch.Tensor:
        dev = self.conv_in.weight.device
        if not torch.is_tensor(timestep):
            timestep = torch.tensor([timestep], dtype=torch.long, device=dev)
        elif timestep.dim() == 0:
            timestep = timestep[None].to(dev)
        timestep = timestep.to(dev) * torch.ones(batch, dtype=timestep.dtype, device=dev)
        return self.time_embedding(timestep_embedding(timestep, BLOCK_OUT_CHANNELS[0]))

    def forward(self, sample, timestep, return_dict: bool = True):
        temb = self.time_embed(timestep, sample.shape[0])
        h = self.conv_in(sample)
        skips = [h]
        for blk in self.down_blocks:
            h, outs = blk(h, temb)
            skips.extend(outs)
        h = self.mid_block(h, temb)
        for blk in self.up_blocks:
            h = blk(h, skips, temb)
        h = self.conv_out(F.silu(self.conv_norm_out(h)))
        return SimpleNamespace(sample=h) if return_dict else (h,)


def build_unet(class_idx: int = 0, perturb_norm: bool = True) -> UNet2DOracle:
    """Random-init weights shared by oracle and CUDA path:
    PyTorch default init under ``manual_seed(1000+class_idx)``, then every GroupNorm
    affine perturbed (gamma=1+0.1 N, beta=0.1 N) so that affine/folding bugs show."""
    g = torch.Generator().manual_seed(1000 + class_idx)
    state = torch.random.get_rng_state()
    torch.manual_seed(1000 + class_idx)
    try:
        m = UNet2DOracle()
    finally:
        torch.random.set_rng_state(state)
    if perturb_norm:
        with torch.no_grad():
            for mod in m.modules():
                if isinstance(mod, nn.GroupNorm):
                    mod.weight.copy_(1.0 + 0.1 * torch.randn(mod.weight.shape, generator=g))
                    mod.bias.copy_(0.1 * torch.randn(mod.bias.shape, generator=g))
    return m.eval()
