"""Drop-in for ``diffusers.UNet2DModel`` as the reference constructs and calls it.

Reference (paths under /root/reference):
  construction     core/generator/model_manager.py:173-194  (same kwargs at image_generator.py:266-288,
                   diffusion/diffusion_generator.py:75-93, xai/XAI.py:313-339)
  weight loading   core/generator/model_manager.py:138-143, xai/XAI.py:604-606 (strict=True)
  call             ``model(latents, t).sample``  core/generator/image_generator.py:400,
                   ``model(latents, timestep=t).sample``  diffusion/diffusion_generator.py:141

The module only HOLDS the parameters (an ``nn.Module`` tree whose ``state_dict()`` keys are
exactly diffusers' -- SURVEY.md A.2); all arithmetic runs in ``libsynt_isic_b200.so``.  There is
no PyTorch or CPU fallback: a forward on a CPU tensor raises.
"""
from __future__ import annotations

import ctypes as C
from types import SimpleNamespace

import numpy as np
import torch
import torch.nn as nn

from . import _lib

SUPPORTED_CONFIG = dict(
    sample_size=128, in_channels=3, out_channels=3, layers_per_block=2,
    block_out_channels=(64, 128, 256, 256),
    down_block_types=("DownBlock2D", "DownBlock2D", "AttnDownBlock2D", "DownBlock2D"),
    up_block_types=("UpBlock2D", "AttnUpBlock2D", "UpBlock2D", "UpBlock2D"),
    class_embed_type=None,
)
DTYPE_CODES = {"fp32": 0, "bf16": 1}


class UNet2DOutput(SimpleNamespace):
    pass


class _Holder(nn.Module):
    """A node of the parameter tree; leaves are ``nn.Parameter``s named weight/bias."""


def _shape_of(name: str, numel: int):
    """diffusers tensor shapes, recovered from the key name and element count."""
    leaf = name.rsplit(".", 1)[1]
    if leaf == "bias" or "norm" in name.split(".")[-2]:
        return (numel,)
    mod = name.rsplit(".", 1)[0]
    if mod.endswith(("conv1", "conv2", "conv_in", "conv_out", "downsamplers.0.conv", "upsamplers.0.conv")):
        return None          # resolved by the caller from channel bookkeeping
    return None


def _init_like_diffusers(name: str, shape, gen: torch.Generator) -> torch.Tensor:
    """PyTorch default layer init (kaiming_uniform(a=sqrt(5)) == U(+-1/sqrt(fan_in)) for weights and
    biases; ones/zeros for norms).  Real use loads a checkpoint over it."""
    leaf = name.rsplit(".", 1)[1]
    parent = name.rsplit(".", 2)[-2]
    if "norm" in parent:
        return torch.ones(shape) if leaf == "weight" else torch.zeros(shape)
    fan_in = int(np.prod(shape[1:])) if len(shape) > 1 else None
    if leaf == "weight":
        bound = 1.0 / (fan_in ** 0.5)
        return (torch.rand(shape, generator=gen) * 2 - 1) * bound
    return torch.zeros(shape)      # bias bound needs the weight's fan_in; set by the caller


class UNet2DModel(nn.Module):
    """``model(sample, timestep).sample`` on B200.

    ``precision``: "bf16" (tcgen05 GEMMs, fp32 accumulate) or "fp32" (verification mode).
    """

    def __init__(self, precision: str = "bf16", **config):
        super().__init__()
        for k, v in config.items():
            if k not in SUPPORTED_CONFIG:
                raise NotImplementedError(f"UNet2DModel argument {k!r} is not part of the reference's configuration")
            want = SUPPORTED_CONFIG[k]
            if (tuple(v) if isinstance(v, (list, tuple)) else v) != want:
                raise NotImplementedError(f"UNet2DModel({k}={v!r}): only the reference's value {want!r} is implemented")
        if precision not in DTYPE_CODES:
            raise ValueError("precision must be 'bf16' or 'fp32'")
        self.config = SimpleNamespace(**SUPPORTED_CONFIG)
        self.precision = precision
        self._handles = {}            # precision -> (handle, version_key)
        self._schedule_key = None
        self._manifest = _lib.unet_manifest()
        gen = torch.Generator().manual_seed(0)
        shapes = _manifest_shapes(self._manifest)
        for name, numel, _ in self._manifest:
            shape = shapes[name]
            assert int(np.prod(shape)) == numel, (name, shape, numel)
            t = _init_like_diffusers(name, shape, gen)
            if name.endswith(".bias") and "norm" not in name.rsplit(".", 2)[-2]:
                wshape = shapes[name[:-4] + "weight"]
                bound = 1.0 / (int(np.prod(wshape[1:])) ** 0.5)
                t = (torch.rand(shape, generator=gen) * 2 - 1) * bound
            self._register(name, t)
        self.eval()

    # ------------------------------------------------------------------ parameter tree
    def _register(self, name: str, value: torch.Tensor):
        parts = name.split(".")
        node = self
        for p in parts[:-1]:
            if p not in node._modules:
                node.add_module(p, _Holder())
            node = node._modules[p]
        node.register_parameter(parts[-1], nn.Parameter(value, requires_grad=False))

    @property
    def device(self):
        return next(self.parameters()).device

    @property
    def dtype(self):
        return torch.float32

    def _version_key(self):
        # the Parameter objects are stable (``.to()`` swaps their ``.data``); walking the module tree on every call cost ~0.2 ms
        plist = self.__dict__.get("_plist")
        if plist is None:
            plist = self.__dict__["_plist"] = list(self.parameters())
        return tuple((p.data_ptr(), p._version) for p in plist)

    def refresh_weights(self):
        """Rebuilds the native weights on the next call.  In-place edits through ``param.data`` (``.data.mul_()``,
        ``.data.copy_()``) bump neither ``data_ptr`` nor ``_version`` and are therefore NOT detected automatically; edits
        through the parameter itself (``load_state_dict``, ``param.mul_()`` under ``no_grad``) are."""
        for h, _ in self._handles.values():
            _lib.lib().synt_unet_destroy(h)
        self._handles = {}
        self._schedule_key = None

    def _packed_params(self) -> np.ndarray:
        sd = self.state_dict()
        total = self._manifest[-1][2] + self._manifest[-1][1]
        blob = np.empty(total, dtype=np.float32)
        for name, numel, off in self._manifest:
            blob[off:off + numel] = sd[name].detach().float().cpu().reshape(-1).numpy()
        return blob

    def _handle(self):
        if self.device.type != "cuda":
            raise RuntimeError("synt_isic_b200.UNet2DModel runs on CUDA (sm_100a) only; move it with .to('cuda')")
        key = self._version_key()
        cur = self._handles.get(self.precision)
        if cur is not None and cur[1] == key:
            return cur[0]
        if cur is not None:
            _lib.lib().synt_unet_destroy(cur[0])
        blob = self._packed_params()
        h = C.c_void_p()
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().synt_unet_create(blob.ctypes.data, blob.size, DTYPE_CODES[self.precision], C.byref(h)),
                       "unet_create")
        self._handles[self.precision] = (h, key)
        self._schedule_key = None
        return h

    def load_state_dict(self, state_dict, strict: bool = True, **kw):
        # pre-0.15 diffusers checkpoints name the attention projections query/key/value/proj_attn
        ren = {".query.": ".to_q.", ".key.": ".to_k.", ".value.": ".to_v.", ".proj_attn.": ".to_out.0."}
        fixed = {}
        for k, v in state_dict.items():
            for a, b in ren.items():
                if ".attentions." in k and a in k:
                    k = k.replace(a, b)
            fixed[k] = v
        return super().load_state_dict(fixed, strict=strict, **kw)

    def __del__(self):
        try:
            for h, _ in self._handles.values():
                _lib.lib().synt_unet_destroy(h)
        except Exception:
            pass

    # ------------------------------------------------------------------ forward -------
    @staticmethod
    def _timestep_int(timestep) -> int:
        if torch.is_tensor(timestep):
            if timestep.numel() != 1:
                u = torch.unique(timestep)
                if u.numel() != 1:
                    raise NotImplementedError("per-sample timesteps are not used by the reference's sampling loops")
                timestep = u
            timestep = timestep.reshape(-1)[0].item()
        t = int(timestep)
        if t != timestep:
            raise NotImplementedError("fractional timesteps are not used by the reference")
        return t

    def forward(self, sample: torch.Tensor, timestep, return_dict: bool = True):
        if not sample.is_cuda:
            raise RuntimeError("synt_isic_b200.UNet2DModel needs a CUDA tensor (no CPU fallback)")
        if sample.dim() != 4 or tuple(sample.shape[1:]) != (3, 128, 128):
            raise ValueError(f"expected [B,3,128,128], got {tuple(sample.shape)}")
        h = self._handle()
        x = sample.contiguous().float()
        out = torch.empty_like(x)
        with torch.cuda.device(x.device):
            _lib.check(_lib.lib().synt_unet_forward(h, x.data_ptr(), x.shape[0], self._timestep_int(timestep),
                                                    out.data_ptr(), _lib.current_stream_ptr()), "unet_forward")
        return UNet2DOutput(sample=out) if return_dict else (out,)

    def debug_tap(self, sample: torch.Tensor, timestep, tap: str) -> torch.Tensor:
        """Output of one internal module (fp32 NCHW) -- parity tests only."""
        h = self._handle()
        x = sample.contiguous().float()
        B = x.shape[0]
        buf = torch.empty(B * 512 * 128 * 128 // 4, dtype=torch.float32, device=x.device)
        c, hh, ww = C.c_int(), C.c_int(), C.c_int()
        with torch.cuda.device(x.device):
            _lib.check(_lib.lib().synt_unet_debug_forward(h, x.data_ptr(), B, self._timestep_int(timestep), tap.encode(),
                                                          buf.data_ptr(), buf.numel(), C.byref(c), C.byref(hh),
                                                          C.byref(ww), _lib.current_stream_ptr()), "unet_debug_forward")
        n = B * c.value * hh.value * ww.value
        return buf[:n].view(B, c.value, hh.value, ww.value).clone()

    # ------------------------------------------------------------------ fused sampling
    def set_schedule(self, scheduler):
        """Uploads ``scheduler.timesteps`` and its per-step coefficients for ``sample``."""
        h = self._handle()
        ts = np.ascontiguousarray(scheduler._timesteps_np, dtype=np.int32)
        coef = np.ascontiguousarray(scheduler._coef, dtype=np.float32)
        key = (ts.tobytes(), coef.tobytes(), self.precision)
        if key != self._schedule_key:
            _lib.check(_lib.lib().synt_unet_set_schedule(h, len(ts), ts.ctypes.data_as(_lib.c_i32p),
                                                         coef.ctypes.data_as(_lib.c_f32p)), "unet_set_schedule")
            self._schedule_key = key

    def sample(self, x: torch.Tensor, scheduler, noise: torch.Tensor | None = None, seed: int = 0,
               image_offset: int = 0, trajectory: torch.Tensor | None = None, eps_tap: torch.Tensor | None = None,
               step_begin: int = 0, step_end: int | None = None, micro_batch: int = 0, use_graph: bool = True,
               step_mask: torch.Tensor | None = None, shared_noise: bool = False,
               image_keys: torch.Tensor | None = None):
        """Runs steps [step_begin, step_end) of the loop at image_generator.py:395-403 IN PLACE on
        ``x`` (fp32 CUDA [B,3,128,128], contiguous).  ``noise`` [n_steps,B,3,128,128] injects z.
        ``step_mask`` (uint8 CUDA [n_steps, B]): image b takes the transition of step s only where the mask is
        non-zero and stays frozen otherwise (coalition decoding of the permutation Time-SHAP); ``shared_noise``:
        all images of the batch draw the same in-kernel noise field.  ``image_keys`` (int64 CUDA [B]): Philox stream id of
        every image (default ``image_offset + b``), so that an image's noise depends on its key only."""
        if not (x.is_cuda and x.is_contiguous() and x.dtype == torch.float32):
            raise ValueError("x must be a contiguous fp32 CUDA tensor")
        self.set_schedule(scheduler)
        n = len(scheduler._timesteps_np)
        step_end = n if step_end is None else step_end
        for t_, nm in ((noise, "noise"), (trajectory, "trajectory"), (eps_tap, "eps_tap")):
            if t_ is not None and not (t_.is_cuda and t_.is_contiguous() and t_.dtype == torch.float32
                                       and tuple(t_.shape) == (n,) + tuple(x.shape)):
                raise ValueError(f"{nm} must be a contiguous fp32 CUDA tensor of shape [n_steps, *x.shape]")
        if step_mask is not None and not (step_mask.is_cuda and step_mask.is_contiguous() and step_mask.dtype == torch.uint8
                                          and tuple(step_mask.shape) == (n, x.shape[0])):
            raise ValueError("step_mask must be a contiguous uint8 CUDA tensor of shape [n_steps, B]")
        if image_keys is not None and not (image_keys.is_cuda and image_keys.is_contiguous() and image_keys.dtype == torch.int64
                                           and tuple(image_keys.shape) == (x.shape[0],)):
            raise ValueError("image_keys must be a contiguous int64 CUDA tensor of shape [B]")
        with torch.cuda.device(x.device):
            masked = step_mask is not None or shared_noise
            if masked:
                _lib.check(_lib.lib().synt_unet_set_step_mask(self._handle(), step_mask.data_ptr() if step_mask is not None else None,
                                                              1 if shared_noise else 0), "unet_set_step_mask")
            if image_keys is not None:
                _lib.check(_lib.lib().synt_unet_set_image_keys(self._handle(), image_keys.data_ptr()), "unet_set_image_keys")
            try:
                self._sample_call(x, noise, seed, image_offset, trajectory, eps_tap, step_begin, step_end, micro_batch, use_graph)
            finally:
                if masked:
                    _lib.lib().synt_unet_set_step_mask(self._handle(), None, 0)
                if image_keys is not None:
                    _lib.lib().synt_unet_set_image_keys(self._handle(), None)
        return x

    def generate_host(self, x_T: np.ndarray, scheduler, seed: int = 0, image_offset: int = 0, micro_batch: int = 0,
                      want_final: bool = False):
        """The host-buffer entry point a non-PyTorch caller binds (``synt_unet_generate_host``): ``x_T`` is a HOST array
        [B,3,128,128] fp32; H2D copy, every step of ``scheduler``, uint8 conversion (image_generator.py:441-447) and the D2H
        copy happen inside the one call.  Returns uint8 [B,128,128,3] (and the fp32 finals with ``want_final``)."""
        self.set_schedule(scheduler)
        x = np.ascontiguousarray(x_T, dtype=np.float32)
        if x.ndim != 4 or x.shape[1:] != (3, 128, 128):
            raise ValueError(f"expected [B,3,128,128], got {x.shape}")
        B = x.shape[0]
        u8 = np.empty((B, 128, 128, 3), np.uint8)
        fin = np.empty_like(x) if want_final else None
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().synt_unet_generate_host(self._handle(), x.ctypes.data, B, int(seed) & 0xFFFFFFFFFFFFFFFF,
                                                          int(image_offset), micro_batch, u8.ctypes.data,
                                                          fin.ctypes.data if fin is not None else None), "unet_generate_host")
        return (u8, fin) if want_final else u8

    def _sample_call(self, x, noise, seed, image_offset, trajectory, eps_tap, step_begin, step_end, micro_batch, use_graph):
        _lib.check(_lib.lib().synt_unet_sample(
            self._handle(), x.data_ptr(), x.shape[0], noise.data_ptr() if noise is not None else None,
            int(seed) & 0xFFFFFFFFFFFFFFFF, int(image_offset),
            trajectory.data_ptr() if trajectory is not None else None,
            eps_tap.data_ptr() if eps_tap is not None else None, step_begin, step_end, micro_batch,
            1 if use_graph else 0, _lib.current_stream_ptr()), "unet_sample")

    PROFILE_CATEGORIES = ("conv_tcgen05", "conv_fp32", "groupnorm_stats", "groupnorm_apply", "attention", "upsample",
                          "conv_in", "conv_out_sched", "misc")

    def profile_step(self, x: torch.Tensor, scheduler, micro_batch: int = 0) -> dict:
        """One eager sampling step with CUDA-event pairs around every kernel (measurement only):
        {category: {"ms", "flops", "launches"}}.  Advances ``x`` by one step."""
        self.set_schedule(scheduler)
        ms = (C.c_double * 16)()
        fl = (C.c_double * 16)()
        ln = (C.c_int * 16)()
        with torch.cuda.device(x.device):
            _lib.check(_lib.lib().synt_unet_profile_step(self._handle(), x.data_ptr(), x.shape[0], micro_batch, ms, fl, ln,
                                                         _lib.current_stream_ptr()), "unet_profile_step")
        return {name: {"ms": ms[i], "flops": fl[i], "launches": ln[i]} for i, name in enumerate(self.PROFILE_CATEGORIES)}

    def profile_records(self):
        """Per-launch rows (category, ms, flops, M, N, K) of the last ``profile_step``."""
        cur = self._handles.get(self.precision)
        if not cur:
            return []
        buf = (C.c_double * (6 * 4096))()
        n = _lib.lib().synt_unet_profile_records(cur[0], buf, 4096)
        return [(self.PROFILE_CATEGORIES[int(buf[6 * i])], buf[6 * i + 1], buf[6 * i + 2], int(buf[6 * i + 3]),
                 int(buf[6 * i + 4]), int(buf[6 * i + 5])) for i in range(n)]

    def launch_count(self) -> int:
        cur = self._handles.get(self.precision)
        return int(_lib.lib().synt_unet_launch_count(cur[0])) if cur else 0

    def workspace_bytes(self) -> int:
        cur = self._handles.get(self.precision)
        return int(_lib.lib().synt_unet_workspace_bytes(cur[0])) if cur else 0


def _manifest_shapes(manifest):
    """Tensor shapes in diffusers layout, derived from key names + element counts."""
    numel = {n: k for n, k, _ in manifest}
    shapes = {}
    for name, k, _ in manifest:
        mod, leaf = name.rsplit(".", 1)
        if leaf == "bias":
            shapes[name] = (k,)
            continue
        out = numel[mod + ".bias"]
        last = mod.rsplit(".", 1)[-1]
        if "norm" in last:
            shapes[name] = (k,)
        elif last in ("conv1", "conv2", "conv_in", "conv_out", "conv"):
            shapes[name] = (out, k // (out * 9), 3, 3)
        elif last == "conv_shortcut":
            shapes[name] = (out, k // out, 1, 1)
        else:                                   # Linear: time_embedding.*, time_emb_proj, to_q/k/v, to_out.0
            shapes[name] = (out, k // out)
    return shapes
