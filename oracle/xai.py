"""TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Restatement of the classifier-evaluation loops of the reference's XAI pipeline, as the
reference issues them (one B=1 forward per call), against any classifier object that
exposes ``get_confidence / get_per_class_score / get_probabilities``:

* ``time_shap``            -- ``ModernXAIAnalyzer.compute_time_shap``  xai/XAI.py:1179-1234
* ``patch_shap``           -- ``compute_shap_approximation``           xai/XAI.py:1111-1177
* ``intervene``            -- ``counterfactual_intervention_advanced`` xai/XAI.py:1454-1597
* ``causal_shift``         -- ``compute_causal_shift_comprehensive``   xai/XAI.py:1600-1700
* ``time_shap_permutation`` -- permutation Shapley over denoising steps, README.md:171-221 (spec only, no reference code)
* ``select_regions``       -- ``select_regions_advanced`` xai/XAI.py:1340-1451 on the REAL numpy ``percentile`` and scipy
  ``ndimage`` calls the reference makes (pinned by construction)
* ``integrated_gradients`` -- ``compute_integrated_gradients`` xai/XAI.py:1039-1085: captum (>=0.6.0, requirements.txt, NOT
  installed here) ``IntegratedGradients.attribute(method='riemann_right')`` restated from its published algorithm
  (Sundararajan et al. 2017 eq. 3 with right Riemann sums: alphas = linspace(1/n, 1, n), step sizes 1/n), gradients from
  torch autograd through the REAL torchvision classifier oracle; ``gradient_attribution`` xai/XAI.py:1087-1109

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


# ------------------------------------------------------------------ permutation Time-SHAP over denoising steps
# README.md:171-221 of the reference specifies it (players = steps, v(S) = logit after a decode that applies the
# transition only on the steps of S, phi_hat = mean over permutations of the marginal contribution of a step to its
# prefix); the reference ships NO code for it, so this oracle is the definition the CUDA path is held to: plain Python
# loops, one B=1 decode per coalition.
def step_permutations(n_steps: int, n_perm: int, seed: int) -> np.ndarray:
    """n_perm permutations of 0..n_steps-1 drawn one after the other from numpy.random.default_rng(seed)."""
    rng = np.random.default_rng(int(seed))
    return np.stack([rng.permutation(n_steps) for _ in range(n_perm)]).astype(np.int64)


def coalition_value(unet, scheduler, classifier, x_T, coalition, noise, target_class) -> float:
    """v(S): decode x_T applying the scheduler transition only on step indices in ``coalition`` (frozen elsewhere),
    same noise realisation ``noise`` [n,1,3,128,128] for every coalition; value = target logit of the classifier."""
    x = x_T.clone()
    with torch.no_grad():
        for i, t in enumerate(scheduler.timesteps):
            if i in coalition:
                eps = unet(x, t).sample
                x = scheduler.step(eps, t, x, noise=noise[i]).prev_sample
        return float(classifier(x)[0, target_class])


def time_shap_permutation(unet, scheduler, classifier, x_T, target_class, n_perm, seed, noise):
    """phi_hat[t] = (1/M) sum_m [ v(Pref_m(t) + {t}) - v(Pref_m(t)) ]; returns (phi [n], prefix values [M, n+1])."""
    n = len(scheduler.timesteps)
    perms = step_permutations(n, n_perm, seed)
    phi = np.zeros(n, np.float64)
    values = np.zeros((n_perm, n + 1), np.float64)
    for m in range(n_perm):
        coalition = set()
        values[m, 0] = coalition_value(unet, scheduler, classifier, x_T, coalition, noise, target_class)
        for k in range(n):
            player = int(perms[m, k])
            coalition = coalition | {player}
            values[m, k + 1] = coalition_value(unet, scheduler, classifier, x_T, coalition, noise, target_class)
            phi[player] += values[m, k + 1] - values[m, k]
    return phi / n_perm, values


def exact_shapley(value_fn, n: int) -> np.ndarray:
    """Brute-force Shapley values of a set function on n players (for checking the permutation estimator)."""
    import itertools
    import math
    phi = np.zeros(n, np.float64)
    for t in range(n):
        others = [p for p in range(n) if p != t]
        for r in range(n):
            for S in itertools.combinations(others, r):
                w = math.factorial(r) * math.factorial(n - r - 1) / math.factorial(n)
                phi[t] += w * (value_fn(set(S) | {t}) - value_fn(set(S)))
    return phi


# ------------------------------------------------------------------ gradient attributions
def gradient_attribution(classifier, image, target_class):
    """XAI.py:1087-1109 -- d get_per_class_score(x, c) / dx by autograd."""
    x = image.detach().clone().requires_grad_(True)
    score = classifier.get_per_class_score(x, target_class)
    score.sum().backward()
    return x.grad.detach().clone()


def integrated_gradients(classifier, image, target_class, baseline, n_steps=50, chunk=10):
    """XAI.py:1039-1085 with the baseline injected.  captum's ``riemann_right`` builder: ``alphas = linspace(1/n, 1, n)``,
    ``step_sizes = [1/n] * n``; ``scaled_features = baseline + alpha * (input - baseline)``; the gradients of the scalar
    forward are multiplied by the step sizes, summed over the path and multiplied by ``input - baseline``.
    Returns (attribution, convergence delta = sum(attr) - (F(x) - F(x')))."""
    alphas = np.linspace(1.0 / n_steps, 1.0, n_steps)
    total = torch.zeros_like(image)
    for i in range(0, n_steps, chunk):
        pts = torch.cat([baseline + float(a) * (image - baseline) for a in alphas[i:i + chunk]])
        total += gradient_attribution(classifier, pts, target_class).sum(0, keepdim=True) * (1.0 / n_steps)
    attr = total * (image - baseline)
    with torch.no_grad():
        delta = attr.sum() - (classifier.get_per_class_score(image, target_class)
                              - classifier.get_per_class_score(baseline, target_class)).sum()
    return attr, float(delta)


def classifier_gradient_taps(classifier, image, target_class):
    """Gradients of the score at the intermediate tensors of the classifier (test diagnostics for the CUDA adjoint chain):
    {'preprocess', 'relu', 'maxpool', 'layerL.J'} -> (activation, d score / d activation), plus 'input'."""
    m = classifier.model
    x = image.detach().clone().requires_grad_(True)
    taps = {}
    z = classifier.preprocess_for_classifier(x); taps["preprocess"] = z
    z = F.relu(m.bn1(m.conv1(z))); taps["relu"] = z
    z = m.maxpool(z); taps["maxpool"] = z
    for l in range(1, 5):
        for j in range(2):
            z = getattr(m, f"layer{l}")[j](z)
            taps[f"layer{l}.{j}"] = z
    logits = m.fc(torch.flatten(m.avgpool(z), 1))
    for t in taps.values():
        t.retain_grad()
    score = torch.log(F.softmax(logits, dim=1)[:, target_class] + 1e-8)
    score.sum().backward()
    out = {k: (t.detach(), t.grad.detach()) for k, t in taps.items()}
    out["input"] = (x.detach(), x.grad.detach())
    return out, score.detach()


# ------------------------------------------------------------------ region selection ----
def select_regions(attribution_map, k_percent=10, region_type="top", morphology_cleanup=True, connectivity=8):
    """XAI.py:1340-1451.  Saliency = channel L2 norm (3-D / 4-D input, batch entry 0) or |x| (2-D); mask = saliency beyond
    the (100-k)-th / k-th numpy percentile; clean-up = closing x2, opening x1 (scipy defaults: outside of the image counts
    as background), then drop connected components smaller than max(10, 1% of the pixels)."""
    from scipy import ndimage
    a = attribution_map.detach().cpu().numpy() if torch.is_tensor(attribution_map) else np.array(attribution_map)
    shape = a.shape
    a = a[0] if a.ndim == 4 else a
    sal = np.linalg.norm(a, axis=0) if a.ndim == 3 else np.abs(a)
    if region_type == "top":
        thr = np.percentile(sal.ravel(), 100 - k_percent)
        mask = sal >= thr
    elif region_type == "bottom":
        thr = np.percentile(sal.ravel(), k_percent)
        mask = sal <= thr
    else:
        raise ValueError(f"unknown region_type {region_type!r}")
    if morphology_cleanup:
        st = ndimage.generate_binary_structure(2, 1 if connectivity == 4 else 2)
        mask = ndimage.binary_opening(ndimage.binary_closing(mask, structure=st, iterations=2), structure=st, iterations=1)
        lab, n = ndimage.label(mask, structure=st)
        if n > 0:
            sizes = ndimage.sum(mask, lab, range(1, n + 1))
            keep = np.where(sizes >= max(10, int(0.01 * mask.size)))[0] + 1
            mask = np.isin(lab, keep)
    sel = sal[mask]
    stats = {"total_pixels": sal.size, "selected_pixels": int(mask.sum()), "target_percentage": k_percent,
             "actual_percentage": mask.sum() / sal.size * 100, "threshold_value": thr,
             "mean_attribution": np.mean(sal), "std_attribution": np.std(sal),
             "mean_attribution_selected": np.mean(sel) if sel.size else 0, "std_attribution_selected": np.std(sel) if sel.size else 0,
             "max_attribution_selected": np.max(sel) if sel.size else 0, "min_attribution_selected": np.min(sel) if sel.size else 0}
    return {"mask": mask, "threshold": thr, "statistics": stats,
            "metadata": {"region_type": region_type, "morphology_cleanup": morphology_cleanup, "connectivity": connectivity,
                         "original_shape": shape}}


# ---------------------------------------------------------------- statistics (XAI.py:1708-2005) ----
def bootstrap_and_permutation(top_k, bottom_k, n_bootstrap=1000, n_permutations=10000):
    """The two resampling loops of ``statistical_validation_comprehensive`` exactly as the reference runs them, drawing from
    the GLOBAL numpy RNG in the reference's call order (bootstrap :1852-1865, then the permutation test :1882-1900).
    Returns (bootstrap_diffs, permuted_diffs) as float64 arrays."""
    top_k = np.array(top_k, dtype=np.float64)
    bottom_k = np.array(bottom_k, dtype=np.float64)
    boot = []
    for _ in range(n_bootstrap):
        top_sample = np.random.choice(top_k, len(top_k), replace=True)
        bottom_sample = np.random.choice(bottom_k, len(bottom_k), replace=True)
        boot.append(np.mean(top_sample) - np.mean(bottom_sample))
    combined = np.concatenate([top_k, bottom_k])
    observed = np.mean(top_k) - np.mean(bottom_k)
    perm = []
    if len(top_k) >= 2 and len(bottom_k) >= 2:
        for _ in range(n_permutations):
            np.random.shuffle(combined)
            perm.append(np.mean(combined[:len(top_k)]) - np.mean(combined[len(top_k):]))
    else:
        perm = [observed]
    return np.array(boot), np.array(perm)


def numpy_pairwise_sum(a):
    """numpy's float64 summation order (DOUBLE_pairwise_sum), restated: what the CUDA replicate means must reproduce."""
    n = len(a)
    if n < 8:
        r = 0.0
        for v in a:
            r += float(v)
        return r
    if n <= 128:
        r = [float(a[j]) for j in range(8)]
        i = 8
        while i < n - (n % 8):
            for j in range(8):
                r[j] += float(a[i + j])
            i += 8
        res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]))
        while i < n:
            res += float(a[i])
            i += 1
        return res
    n2 = n // 2
    n2 -= n2 % 8
    return numpy_pairwise_sum(a[:n2]) + numpy_pairwise_sum(a[n2:])
