"""Drop-in for the reference's ``MelanomaClassifierAdaptive`` (xai/XAI.py:357-471).

Same constructor arguments and methods (``forward``, ``get_probabilities``,
``get_per_class_score``, ``predict``, ``get_confidence``, ``preprocess_for_classifier``); the
``.model`` attribute is a torchvision ``resnet18`` module used ONLY as the parameter container
(so ``state_dict``/``named_parameters`` and the Grad-CAM handle ``.model.layer4[-1].conv2``,
xai/XAI.py:2946, keep working).  Logits are computed by ``libsynt_isic_b200.so``: fused
preprocess kernel, BN-folded tcgen05 implicit-GEMM convolutions, fused pool+FC.  The folded
weights are rebuilt whenever a parameter or BN buffer changes (the reference's sanity check
mutates weights in place, xai/XAI.py:2055-2059).
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F
from torchvision import models

from . import _lib

CLASS_NAMES = ["MEL", "NV", "BCC", "AKIEC", "BKL", "DF", "VASC"]     # xai/XAI.py:196
NUM_CLASSES = 7
CLASSIFIER_IMAGE_SIZE = 224
DTYPE_CODES = {"fp32": 0, "bf16": 1}


class MelanomaClassifierAdaptive(nn.Module):
    def __init__(self, num_classes: int = NUM_CLASSES, architecture: str = "auto", pretrained: bool = True,
                 precision: str = "bf16"):
        super().__init__()
        if num_classes not in (7, 8):
            raise NotImplementedError("the reference builds 7-way (xai_integration.py:79) or 8-way (XAI.py:490) heads")
        self.num_classes = num_classes
        # `pretrained=True` asks torchvision for IMAGENET1K_V1 (XAI.py:389).  The weights come from torchvision's cache or
        # a download; when neither is available (offline box) the network stays random-init and says so LOUDLY -- scores
        # of an untrained classifier are not the reference's.  ``self.pretrained_loaded`` records what happened.
        self.pretrained_loaded = False
        self.model = None
        if pretrained:
            import socket
            old_timeout = socket.getdefaulttimeout()
            try:
                socket.setdefaulttimeout(15.0)                       # an unreachable index must not hang the caller
                self.model = models.resnet18(weights=models.ResNet18_Weights.IMAGENET1K_V1)
                self.pretrained_loaded = True
            except Exception as e:                                   # noqa: BLE001  (no network / no cached file)
                import warnings
                warnings.warn("MelanomaClassifierAdaptive(pretrained=True): the IMAGENET1K_V1 weights could not be loaded "
                              f"({type(e).__name__}: {e}); the classifier is RANDOM-INIT until load_state_dict / "
                              "a classifier checkpoint is applied", RuntimeWarning, stacklevel=2)
            finally:
                socket.setdefaulttimeout(old_timeout)
        if self.model is None:
            self.model = models.resnet18(weights=None)
        self.model.fc = nn.Linear(self.model.fc.in_features, num_classes)
        self.architecture = "resnet18"
        self.precision = precision
        self._handles = {}
        self._manifest = _lib.resnet18_manifest(num_classes)
        self.eval()

    # ------------------------------------------------------------------ handle --------
    def _version_key(self):
        # parameters and BatchNorm buffers, collected once (building the state dict on every call cost ~0.3 ms of a 5 ms
        # Time-SHAP evaluation); ``_apply`` (``.to()``, ``.cuda()``) REPLACES buffer tensors, so it drops the list
        tracked = self.__dict__.get("_tracked")
        if tracked is None:
            sd = self.model.state_dict(keep_vars=True)
            tracked = self.__dict__["_tracked"] = [v for k, v in sd.items() if not k.endswith("num_batches_tracked")]
        return tuple((v.data_ptr(), v._version) for v in tracked)

    def _apply(self, fn, *args, **kwargs):
        self.__dict__.pop("_tracked", None)
        return super()._apply(fn, *args, **kwargs)

    def load_state_dict(self, *args, **kwargs):
        self.__dict__.pop("_tracked", None)                           # (assign=True replaces the tensors)
        return super().load_state_dict(*args, **kwargs)

    def _handle(self):
        dev = next(self.parameters()).device
        if dev.type != "cuda":
            raise RuntimeError("synt_isic_b200 classifier runs on CUDA (sm_100a) only (no CPU fallback)")
        if self.training:
            raise RuntimeError("train-mode BatchNorm is outside the hot path: call .eval() (the reference does)")
        key = self._version_key()
        cur = self._handles.get(self.precision)
        if cur is not None and cur[1] == key:
            return cur[0]
        if cur is not None:
            _lib.lib().synt_resnet18_destroy(cur[0])
        self.__dict__.pop("_tracked", None)                          # weights changed: re-collect (``.model.to()`` replaces buffers)
        key = self._version_key()
        sd = self.model.state_dict()
        total = self._manifest[-1][2] + self._manifest[-1][1]
        blob = np.empty(total, dtype=np.float32)
        for name, numel, off in self._manifest:
            blob[off:off + numel] = sd[name].detach().float().cpu().reshape(-1).numpy()
        h = C.c_void_p()
        with torch.cuda.device(dev):
            _lib.check(_lib.lib().synt_resnet18_create(blob.ctypes.data, blob.size, self.num_classes,
                                                       DTYPE_CODES[self.precision], C.byref(h)), "resnet18_create")
        self._handles[self.precision] = (h, key)
        return h

    def refresh_weights(self):
        """Rebuilds the folded native weights on the next call.  In-place edits through ``param.data`` bump neither
        ``data_ptr`` nor ``_version`` and are NOT detected automatically (edits through the parameter or buffer itself are:
        the reference's sanity check, XAI.py:2055-2059, mutates ``param`` under ``no_grad``)."""
        for h, _ in self._handles.values():
            _lib.lib().synt_resnet18_destroy(h)
        self._handles = {}

    def __del__(self):
        try:
            for h, _ in self._handles.values():
                _lib.lib().synt_resnet18_destroy(h)
        except Exception:
            pass

    # ------------------------------------------------------------------ reference API -
    def preprocess_for_classifier(self, x):
        """Kept for callers that feed ``.model`` directly (Grad-CAM wrapper, XAI.py:2953-2959);
        the fused CUDA kernel inside ``forward`` does the same arithmetic."""
        x = torch.clamp((x + 1.0) / 2.0, 0, 1)
        if x.shape[-1] != CLASSIFIER_IMAGE_SIZE or x.shape[-2] != CLASSIFIER_IMAGE_SIZE:
            x = F.interpolate(x, size=(CLASSIFIER_IMAGE_SIZE, CLASSIFIER_IMAGE_SIZE), mode="bilinear",
                              align_corners=False, antialias=True)
        mean = torch.tensor([0.485, 0.456, 0.406], device=x.device, dtype=x.dtype).view(1, 3, 1, 1)
        std = torch.tensor([0.229, 0.224, 0.225], device=x.device, dtype=x.dtype).view(1, 3, 1, 1)
        return (x - mean) / std

    def forward(self, x):
        dev = next(self.parameters()).device
        if x.device != dev:
            x = x.to(dev)
        if x.dim() != 4 or tuple(x.shape[1:]) != (3, 128, 128):
            raise ValueError(f"expected [B,3,128,128] in [-1,1], got {tuple(x.shape)}")
        h = self._handle()
        x = x.contiguous().float()
        logits = torch.empty(x.shape[0], self.num_classes, dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            _lib.check(_lib.lib().synt_resnet18_logits(h, x.data_ptr(), x.shape[0], logits.data_ptr(),
                                                       _lib.current_stream_ptr()), "resnet18_logits")
        return logits

    def logits_host(self, x: np.ndarray) -> np.ndarray:
        """The host-buffer entry point a non-PyTorch caller binds (``synt_resnet18_logits_host``): HOST array [B,3,128,128]
        fp32 in [-1,1] -> HOST logits [B, num_classes]; the copies happen inside the call."""
        h = self._handle()
        x = np.ascontiguousarray(x, dtype=np.float32)
        if x.ndim != 4 or x.shape[1:] != (3, 128, 128):
            raise ValueError(f"expected [B,3,128,128], got {x.shape}")
        out = np.empty((x.shape[0], self.num_classes), np.float32)
        with torch.cuda.device(next(self.parameters()).device):
            _lib.check(_lib.lib().synt_resnet18_logits_host(h, x.ctypes.data, x.shape[0], out.ctypes.data), "resnet18_logits_host")
        return out

    def get_probabilities(self, x):
        return F.softmax(self.forward(x), dim=1)

    def get_per_class_score(self, x, target_class):
        return torch.log(self.get_probabilities(x)[:, target_class] + 1e-8)

    def predict(self, x):
        with torch.no_grad():
            return torch.argmax(self.forward(x), dim=1)

    def get_confidence(self, x, target_class):
        with torch.no_grad():
            return self.get_probabilities(x)[:, target_class]

    # ------------------------------------------------------------------ input gradient -
    def score_and_input_gradient(self, x, target_class: int):
        """``(s, ds/dx)`` with ``s = get_per_class_score(x, target_class)`` (XAI.py:443-459): the backward pass that
        captum's IntegratedGradients (XAI.py:1039-1085) and ``_compute_gradient_attribution`` (XAI.py:1087-1109) obtain from
        autograd, here as the adjoint chain of the CUDA path (data-gradient convolutions on the same tcgen05 kernels)."""
        dev = next(self.parameters()).device
        if x.device != dev:
            x = x.to(dev)
        if x.dim() != 4 or tuple(x.shape[1:]) != (3, 128, 128):
            raise ValueError(f"expected [B,3,128,128] in [-1,1], got {tuple(x.shape)}")
        h = self._handle()
        x = x.detach().contiguous().float()
        score = torch.empty(x.shape[0], dtype=torch.float32, device=dev)
        grad = torch.empty_like(x)
        with torch.cuda.device(dev):
            _lib.check(_lib.lib().synt_resnet18_score_grad(h, x.data_ptr(), x.shape[0], int(target_class), score.data_ptr(),
                                                           grad.data_ptr(), _lib.current_stream_ptr()), "resnet18_score_grad")
        return score, grad

    def grad_debug_tap(self, x, target_class: int, tap: str):
        h = self._handle()
        x = x.contiguous().float()
        B = x.shape[0]
        buf = torch.empty(B * 64 * 112 * 112, dtype=torch.float32, device=x.device)
        c, hh, ww = C.c_int(), C.c_int(), C.c_int()
        with torch.cuda.device(x.device):
            _lib.check(_lib.lib().synt_resnet18_grad_debug(h, x.data_ptr(), B, int(target_class), tap.encode(), buf.data_ptr(),
                                                           buf.numel(), C.byref(c), C.byref(hh), C.byref(ww),
                                                           _lib.current_stream_ptr()), "resnet18_grad_debug")
        n = B * c.value * hh.value * ww.value
        return buf[:n].view(B, c.value, hh.value, ww.value).clone()

    def debug_tap(self, x, tap: str):
        h = self._handle()
        x = x.contiguous().float()
        B = x.shape[0]
        buf = torch.empty(B * 64 * 112 * 112, dtype=torch.float32, device=x.device)
        c, hh, ww = C.c_int(), C.c_int(), C.c_int()
        with torch.cuda.device(x.device):
            _lib.check(_lib.lib().synt_resnet18_debug(h, x.data_ptr(), B, tap.encode(), buf.data_ptr(), buf.numel(),
                                                      C.byref(c), C.byref(hh), C.byref(ww), _lib.current_stream_ptr()),
                       "resnet18_debug")
        n = B * c.value * hh.value * ww.value
        return buf[:n].view(B, c.value, hh.value, ww.value).clone()

    def profile_forward(self, x) -> dict:
        """Measurement only: device ms of the front end (fused preprocess + stem + max-pool), the body convolutions and
        the pool + FC head of one warm forward of B <= 512 images."""
        h = self._handle()
        x = x.contiguous().float()
        logits = torch.empty(x.shape[0], self.num_classes, dtype=torch.float32, device=x.device)
        ms = (C.c_double * 3)()
        with torch.cuda.device(x.device):
            _lib.check(_lib.lib().synt_resnet18_profile(h, x.data_ptr(), x.shape[0], logits.data_ptr(), ms,
                                                        _lib.current_stream_ptr()), "resnet18_profile")
        return {"front_end_ms": ms[0], "body_ms": ms[1], "head_ms": ms[2]}

    def launch_count(self) -> int:
        cur = self._handles.get(self.precision)
        return int(_lib.lib().synt_resnet18_launch_count(cur[0])) if cur else 0
