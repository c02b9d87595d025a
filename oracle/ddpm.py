"""TEST INFRASTRUCTURE ONLY (see oracle/__init__.py) -- **parity unpinned** vs real diffusers.

fp32 restatement of ``diffusers.DDPMScheduler`` as the reference configures it
(``core/generator/model_manager.py:199-209``: ``num_train_timesteps=1000,
beta_schedule="squaredcos_cap_v2", prediction_type="epsilon"`` then
``set_timesteps(n)``), with diffusers defaults ``variance_type="fixed_small"``,
``clip_sample=True``, ``clip_sample_range=1.0``, ``timestep_spacing="leading"``,
``steps_offset=0``.  The ``"linear"`` schedule used by
``diffusion/diffusion_generator.py:123-128`` is selectable.

All scalar coefficients are computed on fp32 0-d tensors exactly like diffusers does
(``alphas_cumprod`` is a float32 tensor and every derived coefficient is fp32 arithmetic).
"""
from __future__ import annotations

import math
from types import SimpleNamespace

import numpy as np
import torch

NUM_TRAIN_TIMESTEPS = 1000
CLIP_SAMPLE_RANGE = 1.0
VARIANCE_FLOOR = 1e-20


def betas_for_alpha_bar(n: int, max_beta: float = 0.999) -> torch.Tensor:
    """diffusers scheduling_ddpm.py::betas_for_alpha_bar (cosine): python float64 maths,
    then one cast to float32."""
    def alpha_bar(t):
        return math.cos((t + 0.008) / 1.008 * math.pi / 2) ** 2
    betas = []
    for i in range(n):
        t1, t2 = i / n, (i + 1) / n
        betas.append(min(1 - alpha_bar(t2) / alpha_bar(t1), max_beta))
    return torch.tensor(betas, dtype=torch.float32)


class DDPMSchedulerOracle:
    def __init__(self, num_train_timesteps: int = NUM_TRAIN_TIMESTEPS,
                 beta_schedule: str = "squaredcos_cap_v2", beta_start: float = 0.0001,
                 beta_end: float = 0.02, prediction_type: str = "epsilon",
                 clip_sample: bool = True):
        assert prediction_type == "epsilon"
        self.num_train_timesteps = num_train_timesteps
        if beta_schedule == "squaredcos_cap_v2":
            self.betas = betas_for_alpha_bar(num_train_timesteps)
        elif beta_schedule == "linear":
            self.betas = torch.linspace(beta_start, beta_end, num_train_timesteps, dtype=torch.float32)
        else:
            raise NotImplementedError(beta_schedule)
        self.alphas = 1.0 - self.betas
        self.alphas_cumprod = torch.cumprod(self.alphas, dim=0)
        self.one = torch.tensor(1.0)
        self.clip_sample = clip_sample
        self.num_inference_steps = None
        self.timesteps = torch.from_numpy(np.arange(0, num_train_timesteps)[::-1].copy())

    def set_timesteps(self, num_inference_steps: int, device=None):
        if num_inference_steps > self.num_train_timesteps:
            raise ValueError("num_inference_steps > num_train_timesteps")
        self.num_inference_steps = num_inference_steps
        ratio = self.num_train_timesteps // num_inference_steps          # "leading"
        ts = (np.arange(0, num_inference_steps) * ratio).round()[::-1].copy().astype(np.int64)
        self.timesteps = torch.from_numpy(ts).to(device) if device is not None else torch.from_numpy(ts)

    def previous_timestep(self, t: int) -> int:
        if self.num_inference_steps:
            idx = (self.timesteps.cpu() == t).nonzero()[0][0].item()
            return -1 if idx == len(self.timesteps) - 1 else int(self.timesteps[idx + 1])
        return t - 1

    def coefficients(self, t: int):
        """(sqrt(1-abar_t), 1/sqrt(abar_t) as a divisor, c_x0, c_xt, sigma) in fp32."""
        prev_t = self.previous_timestep(t)
        a_t = self.alphas_cumprod[t]
        a_prev = self.alphas_cumprod[prev_t] if prev_t >= 0 else self.one
        b_t = 1 - a_t
        b_prev = 1 - a_prev
        cur_a = a_t / a_prev
        cur_b = 1 - cur_a
        c_x0 = (a_prev ** 0.5 * cur_b) / b_t
        c_xt = cur_a ** 0.5 * b_prev / b_t
        var = torch.clamp((1 - a_prev) / (1 - a_t) * cur_b, min=VARIANCE_FLOOR)
        return b_t ** 0.5, a_t ** 0.5, c_x0, c_xt, var ** 0.5

    def step(self, model_output: torch.Tensor, timestep, sample: torch.Tensor,
             generator=None, noise: torch.Tensor | None = None):
        """diffusers DDPMScheduler.step; ``noise`` injects z (the reference passes no
        generator, image_generator.py:403, so z comes from the global RNG there)."""
        t = int(timestep)
        sqrt_b, sqrt_a, c_x0, c_xt, sigma = self.coefficients(t)
        x0 = (sample - sqrt_b * model_output) / sqrt_a
        if self.clip_sample:
            x0 = x0.clamp(-CLIP_SAMPLE_RANGE, CLIP_SAMPLE_RANGE)
        prev = c_x0 * x0 + c_xt * sample
        if t > 0:
            if noise is None:
                noise = torch.randn(model_output.shape, generator=generator,
                                    device=model_output.device, dtype=model_output.dtype)
            prev = prev + sigma * noise
        return SimpleNamespace(prev_sample=prev, pred_original_sample=x0)

    def add_noise(self, original, noise, timesteps):
        a = self.alphas_cumprod.to(original.device)[timesteps]
        sa = (a ** 0.5).flatten()
        sb = ((1 - a) ** 0.5).flatten()
        while sa.dim() < original.dim():
            sa, sb = sa.unsqueeze(-1), sb.unsqueeze(-1)
        return sa * original + sb * noise
