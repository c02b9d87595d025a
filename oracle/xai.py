"""TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Restatement of the classifier-evaluation loops of the reference's XAI pipeline, as the
reference issues them (one B=1 forward per call), against any classifier object that
exposes ``get_confidence / get_per_class_score / get_probabilities``:

* ``time_shap``            -- ``ModernXAIAnalyzer.compute_time_shap``  xai/XAI.py:1179-1234
* ``patch_shap``           -- ``compute_shap_approximation``           xai/XAI.py:1111-1177
* ``intervene``            -- ``counterfactual_intervention_advanced`` xai/XAI.py:1454-1597
* ``causal_shift``         -- ``compute_causal_shift_comprehensive``   xai/XAI.py:1600-1700

Randomness the reference draws from global RNGs (patch masks ``torch.rand(8,8)``, noise
``randn_like``, ``randperm``) is INJECTED here so the CUDA path can be compared exactly.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F

SHAP_N_SAMPLES = 512        # XAI.py:240
NOISE_STD = 0.5             # XAI.py:262
BLUR_KERNEL_SIZE = 5        # XAI.py:263


def time_shap(classifier, trajectory, timesteps, target_class):
    """XAI.py:1179-1234 -- per frame p_c and log(p_c+1e-8); min-max normalise."""
    conf, prob = [], []
    for image, _t in zip(trajectory, timesteps):
        with torch.no_grad():
            prob.append(classifier.get_confidence(image, target_class).item())
            conf.append(classifier.get_per_class_score(image, target_class).item())
    conf = np.array(conf)
    prob = np.array(prob)
    if len(conf) > 1 and (conf.max() - conf.min()) > 1e-6:
        imp = (conf - conf.min()) / (conf.max() - conf.min())
    else:
        imp = np.ones_like(conf) / len(conf)
    return imp, {"confidence_scores": conf, "probability_scores": prob, "timesteps": timesteps}


def expand_patch_mask(patch_mask: torch.Tensor, patch_size: int = 16) -> torch.Tensor:
    """[nh, nw] bool -> [nh*ps, nw*ps] bool (XAI.py:1151-1157)."""
    return patch_mask.repeat_interleave(patch_size, 0).repeat_interleave(patch_size, 1)


def patch_shap(classifier, image, target_class, patch_masks: torch.Tensor, patch_size: int = 16):
    """XAI.py:1111-1177 with the ``torch.rand(8,8) > 0.5`` masks injected as
    ``patch_masks`` [n_samples, 8, 8] bool."""
    n_samples = patch_masks.shape[0]
    attribution = torch.zeros_like(image)
    with torch.no_grad():
        baseline = classifier.get_per_class_score(torch.zeros_like(image), target_class).item()
    for i in range(n_samples):
        full = expand_patch_mask(patch_masks[i], patch_size)
        masked = image.clone()
        masked[:, :, ~full] = 0
        with torch.no_grad():
            s = classifier.get_per_class_score(masked, target_class).item()
        attribution += (s - baseline) * full[None, None].float()
    return attribution / n_samples


def intervene(image, mask, intervention_type="blur", noise=None, perm_seed=None,
              noise_std=NOISE_STD, blur_kernel=BLUR_KERNEL_SIZE):
    """XAI.py:1454-1597: x~ = clamp(x (1-M) + I M, -1, 1).  ``noise`` injects the
    ``randn_like`` tensor of the 'noise'/'gaussian_noise' types."""
    m = torch.as_tensor(mask).float().to(image.device)
    while m.dim() < image.dim():
        m = m.unsqueeze(0)
    m = m.expand_as(image)
    if intervention_type == "noise":
        inter = noise * noise_std
    elif intervention_type == "gaussian_noise":
        inter = noise * max(noise_std, image.std().item() * 0.5)
    elif intervention_type == "zero":
        inter = torch.zeros_like(image)
    elif intervention_type == "mean":
        inter = torch.zeros_like(image) + image.mean(dim=[-2, -1], keepdim=True)
    elif intervention_type in ("blur", "inpaint"):
        k = blur_kernel + (1 - blur_kernel % 2)
        inter = F.avg_pool2d(image, kernel_size=k, stride=1, padding=k // 2)
    elif intervention_type == "shuffle":
        inter = image.clone()
        g = torch.Generator().manual_seed(perm_seed or 0)
        for b in range(image.shape[0]):
            for c in range(image.shape[1]):
                sel = m[b, c].bool()
                px = inter[b, c][sel]
                if len(px) > 1:
                    inter[b, c][sel] = px[torch.randperm(len(px), generator=g)]
    else:
        raise ValueError(intervention_type)
    return torch.clamp(image * (1 - m) + inter * m, -1, 1), inter


def causal_shift(classifier, original, modified, target_class):
    """XAI.py:1600-1700 (scalar outputs only).  The reference issues 18 forwards for the
    2 distinct inputs; the values depend only on the 2 probability vectors."""
    with torch.no_grad():
        po = classifier.get_probabilities(original)
        pm = classifier.get_probabilities(modified)
    so = torch.log(po[:, target_class] + 1e-8)
    sm = torch.log(pm[:, target_class] + 1e-8)
    cfi = so - sm
    delta = cfi.abs() / (so.abs() + 1e-8)
    all_so, all_sm = torch.log(po[0] + 1e-8), torch.log(pm[0] + 1e-8)
    mid = torch.log((po + pm) / 2 + 1e-8)
    return {
        "cfi": float(cfi), "delta": float(delta),
        "original_score": float(so), "modified_score": float(sm),
        "original_probability": float(po[0, target_class]),
        "modified_probability": float(pm[0, target_class]),
        "probability_shift": float(po[0, target_class] - pm[0, target_class]),
        "original_prediction": int(po.argmax(1)[0]), "modified_prediction": int(pm.argmax(1)[0]),
        "prediction_changed": bool(po.argmax(1)[0] != pm.argmax(1)[0]),
        "confidence_drop": float(po.max() - pm.max()),
        "all_cfi": (all_so - all_sm).tolist(),
        "all_delta": ((all_so - all_sm).abs() / (all_so.abs() + 1e-8)).tolist(),
        "kl_divergence": float(F.kl_div(torch.log(pm + 1e-8), po, reduction="sum")),
        "js_divergence": float(0.5 * (F.kl_div(mid, po, reduction="sum") + F.kl_div(mid, pm, reduction="sum"))),
        "total_variation": float(0.5 * (po - pm).abs().sum()),
    }
