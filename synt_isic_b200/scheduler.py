"""Drop-in for ``diffusers.DDPMScheduler`` as the reference uses it.

Reference call sites (paths under /root/reference):
  construction       core/generator/model_manager.py:199-203, core/generator/image_generator.py:292-296,
                     diffusion/diffusion_generator.py:123-128 ("linear" schedule)
  set_timesteps(n)   core/generator/model_manager.py:209, xai/XAI.py:739
  .timesteps         core/generator/image_generator.py:395 (iterated: 0-d int64 tensors, descending)
  .step(eps, t, x).prev_sample   core/generator/image_generator.py:403
  .add_noise         diffusion/train_diffusion.py:233

Tables (betas, alphas_cumprod, timesteps, per-step coefficients) come from the library's
host-side ``synt_ddpm_tables``; ``step`` launches the CUDA kernel ``synt_ddpm_step``.  CPU
tensors are rejected: this package has no CPU path.
"""
from __future__ import annotations

import ctypes as C
import math
from types import SimpleNamespace

import numpy as np
import torch

from . import _lib

_SCHEDULES = {"squaredcos_cap_v2": 0, "linear": 1}


class DDPMSchedulerOutput(SimpleNamespace):
    pass


class DDPMScheduler:
    order = 1

    def __init__(self, num_train_timesteps: int = 1000, beta_start: float = 0.0001, beta_end: float = 0.02,
                 beta_schedule: str = "linear", prediction_type: str = "epsilon", variance_type: str = "fixed_small",
                 clip_sample: bool = True, clip_sample_range: float = 1.0, timestep_spacing: str = "leading",
                 steps_offset: int = 0, **unused):
        if beta_schedule not in _SCHEDULES:
            raise NotImplementedError(f"beta_schedule={beta_schedule!r} (reference uses squaredcos_cap_v2 / linear)")
        if prediction_type != "epsilon" or variance_type != "fixed_small" or not clip_sample \
                or clip_sample_range != 1.0 or timestep_spacing != "leading" or steps_offset != 0:
            raise NotImplementedError("only the reference's DDPMScheduler configuration is implemented")
        self.config = SimpleNamespace(num_train_timesteps=num_train_timesteps, beta_start=beta_start,
                                      beta_end=beta_end, beta_schedule=beta_schedule,
                                      prediction_type=prediction_type, variance_type=variance_type,
                                      clip_sample=clip_sample, clip_sample_range=clip_sample_range,
                                      timestep_spacing=timestep_spacing, steps_offset=steps_offset)
        self.init_noise_sigma = 1.0
        self.num_inference_steps = None
        self._build(num_train_timesteps)          # diffusers default: every train timestep, descending
        self.num_inference_steps = None

    # ------------------------------------------------------------------ tables --------
    def _build(self, n_steps: int, device=None):
        """Host-side tables, computed with the same fp32 torch scalar arithmetic diffusers uses
        (CPU ``torch.cumprod`` accumulates in double; ``** 0.5`` is torch's vectorised sqrt), so the
        coefficients are bit-identical to the reference's.  ``synt_ddpm_tables`` is the C twin for
        non-Python callers (timesteps identical, coefficients within 1 ulp)."""
        cfg = self.config
        n_train = cfg.num_train_timesteps
        if cfg.beta_schedule == "squaredcos_cap_v2":
            def alpha_bar(t):
                return math.cos((t + 0.008) / 1.008 * math.pi / 2) ** 2
            betas = torch.tensor([min(1 - alpha_bar((i + 1) / n_train) / alpha_bar(i / n_train), 0.999)
                                  for i in range(n_train)], dtype=torch.float32)
        else:
            betas = torch.linspace(cfg.beta_start, cfg.beta_end, n_train, dtype=torch.float32)
        self.betas = betas
        self.alphas = 1.0 - betas
        self.alphas_cumprod = torch.cumprod(self.alphas, dim=0)
        ratio = n_train // n_steps                                     # "leading" spacing
        ts = (np.arange(0, n_steps) * ratio).round()[::-1].copy().astype(np.int64)
        one = torch.tensor(1.0)
        coef = np.zeros((n_steps, 5), dtype=np.float32)
        for i, t in enumerate(ts.tolist()):
            prev_t = int(ts[i + 1]) if i + 1 < n_steps else -1
            a_t = self.alphas_cumprod[t]
            a_prev = self.alphas_cumprod[prev_t] if prev_t >= 0 else one
            b_t, b_prev = 1 - a_t, 1 - a_prev
            cur_a = a_t / a_prev
            cur_b = 1 - cur_a
            var = torch.clamp((1 - a_prev) / (1 - a_t) * cur_b, min=1e-20)
            coef[i] = [float(b_t ** 0.5), float(a_t ** 0.5), float((a_prev ** 0.5 * cur_b) / b_t),
                       float(cur_a ** 0.5 * b_prev / b_t), float(var ** 0.5) if t > 0 else 0.0]
        self._timesteps_np = ts
        self._coef = coef                         # [n][5] = sqrt(1-abar_t), sqrt(abar_t), c_x0, c_xt, sigma
        self.timesteps = torch.from_numpy(ts.copy())
        if device is not None:
            self.timesteps = self.timesteps.to(device)
        self._index = {int(t): i for i, t in enumerate(ts)}

    @staticmethod
    def c_tables(n_train: int, schedule: str, beta_start: float, beta_end: float, n_steps: int):
        """The library's own host-side tables (``synt_ddpm_tables``)."""
        ts = np.zeros(n_steps, dtype=np.int64)
        coef = np.zeros((n_steps, 5), dtype=np.float32)
        acp = np.zeros(n_train, dtype=np.float32)
        _lib.check(_lib.lib().synt_ddpm_tables(
            n_train, _SCHEDULES[schedule], beta_start, beta_end, n_steps,
            ts.ctypes.data_as(_lib.c_i64p), coef.ctypes.data_as(_lib.c_f32p), acp.ctypes.data_as(_lib.c_f32p)),
            "ddpm_tables")
        return ts, coef, acp

    def set_timesteps(self, num_inference_steps: int, device=None):
        if num_inference_steps > self.config.num_train_timesteps:
            raise ValueError("`num_inference_steps` cannot be larger than `num_train_timesteps`")
        self._build(int(num_inference_steps), device)
        self.num_inference_steps = int(num_inference_steps)

    def scale_model_input(self, sample, timestep=None):
        return sample

    def __len__(self):
        return self.config.num_train_timesteps

    def coefficients(self, timestep) -> np.ndarray:
        t = int(timestep)
        if t not in self._index:
            raise ValueError(f"timestep {t} is not on the current schedule")
        return self._coef[self._index[t]]

    # ------------------------------------------------------------------ step ----------
    def step(self, model_output: torch.Tensor, timestep, sample: torch.Tensor, generator=None,
             return_dict: bool = True, noise: torch.Tensor | None = None):
        if not (model_output.is_cuda and sample.is_cuda):
            raise RuntimeError("synt_isic_b200.DDPMScheduler.step needs CUDA tensors (no CPU fallback)")
        c = np.ascontiguousarray(self.coefficients(timestep), dtype=np.float32)
        eps = model_output.contiguous().float()
        x = sample.contiguous().float()
        z = None
        if c[4] != 0.0:
            # the reference passes no generator (image_generator.py:403): global device RNG
            z = noise if noise is not None else torch.randn(eps.shape, generator=generator, device=eps.device,
                                                            dtype=eps.dtype)
            z = z.contiguous().float()
        out = torch.empty_like(x)
        with torch.cuda.device(x.device):                       # the tensor's device, not whichever happens to be current
            _lib.check(_lib.lib().synt_ddpm_step(eps.data_ptr(), x.data_ptr(), z.data_ptr() if z is not None else None,
                                                 out.data_ptr(), x.numel(), c.ctypes.data_as(_lib.c_f32p),
                                                 _lib.current_stream_ptr()), "ddpm_step")
        if not return_dict:
            return (out,)
        return DDPMSchedulerOutput(prev_sample=out)

    def add_noise(self, original_samples, noise, timesteps):
        a = self.alphas_cumprod.to(original_samples.device)[timesteps]
        sa, sb = (a ** 0.5).flatten(), ((1 - a) ** 0.5).flatten()
        while sa.dim() < original_samples.dim():
            sa, sb = sa.unsqueeze(-1), sb.unsqueeze(-1)
        return sa * original_samples + sb * noise
