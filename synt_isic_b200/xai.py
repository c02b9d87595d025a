"""Classifier-evaluation loops of the reference's XAI pipeline, batched for B200.

Every function keeps the reference's name, arguments and result structure; what changes is how
the ResNet18 evaluations are issued: the reference calls the classifier at batch 1 once (or
twice, or 18 times) per frame / coalition / intervention with a ``.item()`` sync each
(SURVEY.md 3.2); here the distinct inputs are built on the GPU, evaluated in ONE batched call
of the CUDA classifier and (optionally) sharded across ranks with a single all_gather.

  compute_time_shap                     xai/XAI.py:1179-1234   (2 forwards/frame -> 1 batched)
  compute_shap_approximation            xai/XAI.py:1111-1177   (513 forwards -> 1 batched)
  counterfactual_intervention_advanced  xai/XAI.py:1454-1597
  compute_causal_shift_comprehensive    xai/XAI.py:1600-1700   (18 forwards -> 1 batch of 2)
  compute_integrated_gradients          xai/XAI.py:1039-1085   (captum autograd -> batched CUDA adjoint)
  compute_integrated_gradients_batch    the same for every frame of a trajectory (pipeline stage 1, xai/XAI.py:2740-2751)
  compute_gradient_attribution          xai/XAI.py:1087-1109
  compute_combined_attribution          xai/XAI.py:1236-1291
  select_regions_advanced               xai/XAI.py:1340-1451   (numpy percentile + scipy.ndimage -> one CTA per map)
  statistical_validation_comprehensive  xai/XAI.py:1708-2005   (bootstrap / permutation loops -> one thread per replicate)
  IntegratedXAIAnalyzer                 xai/xai_integration.py:75-132

Out of scope: Grad-CAM and plots.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F

from . import _lib
from .classifier import CLASS_NAMES, MelanomaClassifierAdaptive
from .dist import _world, gather_rows, shard_bounds, sharded_eval

SHAP_N_SAMPLES = 512       # xai/XAI.py:240
NOISE_STD = 0.5            # xai/XAI.py:262
BLUR_KERNEL_SIZE = 5       # xai/XAI.py:263
TOP_K_PERCENT = 10         # xai/XAI.py:238
BOTTOM_K_PERCENT = 10      # xai/XAI.py:239
IG_N_STEPS = 50            # xai/XAI.py:240
_TYPE_CODES = {"zero": 0, "mean": 1, "blur": 2, "inpaint": 2, "noise": 3, "gaussian_noise": 3, "shuffle": 4}


def _dev(classifier):
    return next(classifier.parameters()).device


def _probs(classifier, images: torch.Tensor, group=None) -> torch.Tensor:
    """softmax(logits) for a batch of [-1,1] images; the batch is split across ranks when a
    process group is given (one all_gather of the [n,7] logits)."""
    with torch.no_grad():
        logits = sharded_eval(classifier, images, group)
    return F.softmax(logits, dim=1)


def _probs_from_host(classifier, frames_cpu: torch.Tensor, dev, chunk: int = 256, group=None) -> torch.Tensor:
    """softmax(logits) of frames that live in (ideally pinned) HOST memory: the frames are copied in chunks on a side
    stream so that the H2D copy of chunk i+1 overlaps the classifier kernels of chunk i.  With ``group`` every rank uploads
    and evaluates only its contiguous slice of the frames; one all_gather of the logits."""
    n_all = frames_cpu.shape[0]
    rank, world = _world(group)
    if world > 1:
        lo, hi = shard_bounds(n_all, rank, world)
        local = _probs_from_host(classifier, frames_cpu[lo:hi], dev, chunk) if hi > lo else None   # softmax is row-wise
        return gather_rows(local, n_all, group, dev)
    n = frames_cpu.shape[0]
    main = torch.cuda.current_stream(dev)
    side = torch.cuda.Stream(device=dev)
    bufs = [torch.empty((min(chunk, n), 3, 128, 128), dtype=torch.float32, device=dev) for _ in range(2)]
    ready = [torch.cuda.Event() for _ in range(2)]
    used = [torch.cuda.Event() for _ in range(2)]
    out = []
    side.wait_stream(main)

    def issue(i):
        b, lo = i & 1, i * chunk
        with torch.cuda.stream(side):
            side.wait_event(used[b])
            bufs[b][: min(chunk, n - lo)].copy_(frames_cpu[lo: lo + chunk], non_blocking=True)
            ready[b].record(side)

    n_chunks = (n + chunk - 1) // chunk
    issue(0)
    with torch.no_grad():
        for i in range(n_chunks):
            if i + 1 < n_chunks:
                issue(i + 1)
            b, lo = i & 1, i * chunk
            main.wait_event(ready[b])
            out.append(classifier(bufs[b][: min(chunk, n - lo)]))
            used[b].record(main)
    main.wait_stream(side)
    return F.softmax(torch.cat(out), dim=1)


# ------------------------------------------------------------------ Time-SHAP -----------
def compute_time_shap(classifier, trajectory, timesteps, target_class, group=None, verbose=False):
    """xai/XAI.py:1179-1234.  ``trajectory``: list of [1,3,128,128] tensors or one [T,3,128,128]."""
    dev = _dev(classifier)
    frames = trajectory if torch.is_tensor(trajectory) else torch.cat([f.to(dev).reshape(-1, 3, 128, 128) for f in trajectory])
    if frames.device.type == "cpu" and dev.type == "cuda" and frames.shape[0] > 256:
        # a host-resident trajectory: stream it through the classifier (copies overlap the kernels); with a process group
        # every rank uploads only its slice
        p = _probs_from_host(classifier, frames.reshape(-1, 3, 128, 128).float(), dev, group=group)[:, target_class]
    else:
        frames = frames.to(dev).reshape(-1, 3, 128, 128)
        p = _probs(classifier, frames, group)[:, target_class]
    both = torch.stack([p, torch.log(p + 1e-8)]).double().cpu().numpy()   # ONE device-to-host copy (and one sync) for both score rows
    prob_scores = both[0]                                       # get_confidence(...).item()
    confidence_scores = both[1]                                 # get_per_class_score(...).item()
    if len(confidence_scores) > 1 and (confidence_scores.max() - confidence_scores.min()) > 1e-6:
        imp = (confidence_scores - confidence_scores.min()) / (confidence_scores.max() - confidence_scores.min())
    else:
        imp = np.ones_like(confidence_scores) / len(confidence_scores)
    raw = {"confidence_scores": confidence_scores, "probability_scores": prob_scores, "timesteps": timesteps}
    if verbose:
        k = int(np.argmax(imp))
        print(f"Time-SHAP: most important step t={timesteps[k]} (importance {imp[k]:.3f})")
    return imp, raw


# ------------------------------------------------------------------ permutation Time-SHAP
def draw_step_permutations(n_steps: int, n_perm: int, seed: int) -> np.ndarray:
    """The coalition enumeration: ``n_perm`` permutations of the step indices 0..n_steps-1 (index into
    ``scheduler.timesteps``, i.e. 0 = the noisiest step), drawn one after the other from
    ``numpy.random.default_rng(seed)``.  Deterministic in (n_steps, n_perm, seed); int64 [n_perm, n_steps]."""
    rng = np.random.default_rng(int(seed))
    return np.stack([rng.permutation(n_steps) for _ in range(n_perm)]).astype(np.int64)


def prefix_coalitions(perm: np.ndarray) -> np.ndarray:
    """Prefix coalitions of one permutation as a mask [n_steps + 1, n_steps]: row k holds the first k players."""
    n = len(perm)
    m = np.zeros((n + 1, n), np.uint8)
    for k in range(n):
        m[k + 1] = m[k]
        m[k + 1, perm[k]] = 1
    return m


def shapley_from_prefix_values(perms: np.ndarray, values: np.ndarray) -> np.ndarray:
    """phi_hat[t] = mean over permutations of v(Pref(t) + {t}) - v(Pref(t)) (README.md:199-207 of the reference).
    ``values`` [n_perm, n_steps + 1]: value of every prefix coalition, float64."""
    n_perm, n = perms.shape
    phi = np.zeros(n, np.float64)
    for m in range(n_perm):
        phi[perms[m]] += values[m, 1:] - values[m, :-1]
    return phi / n_perm


def compute_time_shap_permutation(model, scheduler, classifier, x_T: torch.Tensor, target_class: int, n_perm: int = 8,
                                  seed: int = 0, noise: torch.Tensor | None = None, noise_seed: int = 0, group=None):
    """Permutation estimate of the Shapley value of every denoising step (README.md:171-221 of the reference; the
    reference ships no code for it, so the enumeration order is defined here: ``draw_step_permutations``).

    Players: the ``n`` steps of ``scheduler.timesteps``.  v(S) = logit_target(classifier(Dec(x_T; S))) where Dec runs the
    sampling loop and applies the transition only at steps in S (the image is frozen on the others); all coalitions share
    one noise realisation (``noise`` [n,1,3,128,128] injected, else the in-kernel Philox field of ``noise_seed``).
    Per permutation the n+1 prefix coalitions are decoded as ONE batch of the CUDA sampler (per-image step mask) and
    scored in one classifier call; with ``group`` the permutations are split over the ranks and the partial sums are
    combined by one all_reduce.  Returns (phi [n] float64, raw dict)."""
    import torch.distributed as dist
    dev = x_T.device
    n = len(scheduler.timesteps)
    perms = draw_step_permutations(n, n_perm, seed)
    rank, world = (dist.get_rank(group), dist.get_world_size(group)) if (group is not None and dist.is_initialized()) else (0, 1)
    values = np.zeros((n_perm, n + 1), np.float64)
    B = n + 1
    z = noise.to(dev).float().expand(n, B, 3, 128, 128).contiguous() if noise is not None else None
    for m in range(rank, n_perm, world):
        mask = torch.from_numpy(np.ascontiguousarray(prefix_coalitions(perms[m]).T)).to(dev)      # [n_steps, B]
        x = x_T.to(dev).float().reshape(1, 3, 128, 128).repeat(B, 1, 1, 1).contiguous()
        model.sample(x, scheduler, noise=z, seed=noise_seed, step_mask=mask, shared_noise=True)
        with torch.no_grad():
            logits = classifier(x)
        values[m] = logits[:, target_class].double().cpu().numpy()
    if world > 1:
        t = torch.from_numpy(values).to(dev)
        dist.all_reduce(t, group=group)                                  # rows of other ranks are zero here
        values = t.cpu().numpy()
    phi = shapley_from_prefix_values(perms, values)
    raw = {"permutations": perms, "prefix_values": values, "v_empty": float(values[0, 0]), "v_full": float(values[0, -1]),
           "efficiency_gap": float(phi.sum() - (values[:, -1] - values[:, 0]).mean())}
    return phi, raw


# ------------------------------------------------------------------ patch-SHAP ----------
def draw_patch_masks(n_samples: int, n_h: int = 8, n_w: int = 8) -> torch.Tensor:
    """The coalition masks exactly as the reference draws them: one ``torch.rand(n_h, n_w) > 0.5``
    per sample from the global CPU RNG (xai/XAI.py:1145)."""
    return torch.stack([torch.rand(n_h, n_w) > 0.5 for _ in range(n_samples)])


def compute_shap_approximation(classifier, image, target_class, n_samples=SHAP_N_SAMPLES, patch_size=16,
                               patch_masks: torch.Tensor | None = None, group=None):
    """xai/XAI.py:1111-1177.  ``patch_masks`` [n,H/p,W/p] bool injects the coalitions."""
    dev = _dev(classifier)
    image = image.to(dev).float()
    _, ch, height, width = image.shape
    n_h, n_w = height // patch_size, width // patch_size
    if patch_masks is None:
        patch_masks = draw_patch_masks(n_samples, n_h, n_w)
    n_samples = patch_masks.shape[0]
    pm = patch_masks.to(dev).to(torch.uint8).contiguous()
    batch = torch.empty(n_samples + 1, ch, height, width, dtype=torch.float32, device=dev)
    batch[0].zero_()                                            # baseline: black image in [-1,1] space
    x0 = image[0].contiguous()
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().synt_patch_mask_apply(x0.data_ptr(), pm.data_ptr(), n_samples, ch, height, width,
                                                    patch_size, batch[1:].data_ptr(), _lib.current_stream_ptr()),
                   "patch_mask_apply")
    scores = torch.log(_probs(classifier, batch, group)[:, target_class] + 1e-8)
    contrib = scores[1:] - scores[0]                            # masked_score - baseline_score
    patch_attr = (contrib[:, None, None] * pm.float()).sum(0) / n_samples
    full = patch_attr.repeat_interleave(patch_size, 0).repeat_interleave(patch_size, 1)
    return full[None, None].expand(1, ch, height, width).contiguous()


def compute_shap_approximation_batch(classifier, images, target_class, n_samples=SHAP_N_SAMPLES, patch_size=16,
                                     patch_masks: torch.Tensor | None = None, frames_per_pass: int = 8, group=None):
    """``compute_shap_approximation`` for a stack of frames [T,3,128,128] -- what stage 1 of the reference pipeline runs frame
    by frame over the WHOLE trajectory (XAI.py:2745-2747).  ``patch_masks`` [T,n,H/p,W/p] injects the coalitions; otherwise
    they are drawn frame after frame from the global CPU RNG in the reference's order (XAI.py:1145).  ``frames_per_pass``
    frames x (n + 1) masked images go through ONE classifier call.  Returns [T,3,128,128]."""
    dev = _dev(classifier)
    x = images.to(dev).float().reshape(-1, 3, 128, 128).contiguous()
    T, ch, height, width = x.shape
    n_h, n_w = height // patch_size, width // patch_size
    if patch_masks is None:
        patch_masks = torch.stack([draw_patch_masks(n_samples, n_h, n_w) for _ in range(T)])
    n_samples = patch_masks.shape[1]
    out = torch.empty_like(x)
    L = _lib.lib()
    per = n_samples + 1
    for t0 in range(0, T, frames_per_pass):
        g = min(frames_per_pass, T - t0)
        pm = patch_masks[t0:t0 + g].to(dev).to(torch.uint8).contiguous()
        batch = torch.empty(g * per, ch, height, width, dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            st = _lib.current_stream_ptr()
            for j in range(g):
                batch[j * per].zero_()                              # baseline: black image in [-1,1] space
                _lib.check(L.synt_patch_mask_apply(x[t0 + j].data_ptr(), pm[j].data_ptr(), n_samples, ch, height, width,
                                                   patch_size, batch[j * per + 1:(j + 1) * per].data_ptr(), st), "patch_mask_apply")
        scores = torch.log(_probs(classifier, batch, group)[:, target_class] + 1e-8).view(g, per)
        contrib = scores[:, 1:] - scores[:, :1]                    # masked_score - baseline_score
        patch_attr = (contrib[:, :, None, None] * pm.float()).sum(1) / n_samples
        full = patch_attr.repeat_interleave(patch_size, 1).repeat_interleave(patch_size, 2)
        out[t0:t0 + g] = full[:, None].expand(g, ch, height, width)
    return out


# ------------------------------------------------------------------ gradient attributions
def get_baseline(image: torch.Tensor, baseline_type: str = "noise", generator=None) -> torch.Tensor:
    """xai/XAI.py:1008-1037: 'noise' = 0.1 N(0,1), 'blur' = 31x31 box blur, anything else zeros.  (The reference caches
    one baseline per (type, shape, device) for the life of the analyzer; pass ``baseline=`` to the callers for that.)"""
    if baseline_type == "noise":
        return torch.randn(image.shape, device=image.device, dtype=image.dtype, generator=generator) * 0.1
    if baseline_type == "blur":
        return F.avg_pool2d(image, kernel_size=31, stride=1, padding=15)
    return torch.zeros_like(image)


def compute_gradient_attribution(classifier, image, target_class):
    """xai/XAI.py:1087-1109: d get_per_class_score(x, c) / dx, one call of the CUDA adjoint chain."""
    dev = _dev(classifier)
    x = image.to(dev).float().reshape(-1, 3, 128, 128)
    return classifier.score_and_input_gradient(x, target_class)[1]


def compute_integrated_gradients(classifier, image, target_class, n_steps=IG_N_STEPS, baseline_type="noise",
                                 baseline=None, generator=None, return_convergence_delta=False):
    """xai/XAI.py:1039-1085 = captum ``IntegratedGradients(forward).attribute(x, baselines=x', n_steps=n,
    method='riemann_right')`` with forward = ``get_per_class_score(., c)``:

        IG = (x - x') * sum_{k=1..n} (1/n) * ds/dx (x' + (k/n) (x - x'))

    The reference runs the n path points through autograd (captum's default internal batching); here they are ONE batch
    of the CUDA classifier's score + input-gradient pass.  ``baseline`` injects x' (tests; the reference caches its
    random baseline).  With ``return_convergence_delta`` also returns sum(IG) - (s(x) - s(x')) (captum's delta)."""
    dev = _dev(classifier)
    x = image.to(dev).float().reshape(1, 3, 128, 128).contiguous()
    base = (get_baseline(x, baseline_type, generator) if baseline is None else baseline.to(dev).float()).reshape(1, 3, 128, 128).contiguous()
    per = x.numel()
    pts = torch.empty(n_steps, 3, 128, 128, dtype=torch.float32, device=dev)
    out = torch.empty_like(x)
    with torch.cuda.device(dev):
        st = _lib.current_stream_ptr()
        _lib.check(_lib.lib().synt_ig_interpolate(x.data_ptr(), base.data_ptr(), n_steps, per, pts.data_ptr(), st), "ig_interpolate")
        _, grads = classifier.score_and_input_gradient(pts, target_class)
        _lib.check(_lib.lib().synt_ig_reduce(grads.data_ptr(), x.data_ptr(), base.data_ptr(), n_steps, per, out.data_ptr(), st), "ig_reduce")
    if return_convergence_delta:
        s = classifier.get_per_class_score(torch.cat([x, base]), target_class)
        return out, float(out.sum() - (s[0] - s[1]))
    return out


def compute_integrated_gradients_batch(classifier, images, target_class, n_steps=IG_N_STEPS, baseline_type="noise",
                                        baselines=None, generator=None, images_per_pass: int = 8, group=None):
    """Integrated Gradients for a stack of images [n,3,128,128] -- what stage 1 of the reference pipeline computes frame by
    frame for the WHOLE trajectory (XAI.py:2740-2751).  ``images_per_pass * n_steps`` path points go through one score +
    input-gradient call.  ``baselines``: [1,3,128,128] shared by all images (the reference caches ONE baseline per shape,
    XAI.py:1021-1037) or [n,3,128,128]; drawn once with ``baseline_type`` when omitted.  ``group``: split the images over
    the ranks of a torch.distributed group (one all_gather).  Returns [n,3,128,128]."""
    dev = _dev(classifier)
    x = images.to(dev).float().reshape(-1, 3, 128, 128).contiguous()
    n, per = x.shape[0], 3 * 128 * 128
    if baselines is None:
        baselines = get_baseline(x[:1], baseline_type, generator)
    base = baselines.to(dev).float().reshape(-1, 3, 128, 128).contiguous()
    if base.shape[0] not in (1, n):
        raise ValueError(f"baselines must have 1 or {n} entries, got {base.shape[0]}")
    if group is not None:
        # unit = one image: contiguous split over the ranks (weights replicated), ONE all_gather of the maps.  Every rank
        # must pass the same images AND baselines (draw a random baseline once and broadcast it, or inject it).
        both = torch.cat([x, base.expand(n, -1, -1, -1)], dim=1)
        flat = sharded_eval(lambda t: compute_integrated_gradients_batch(
            classifier, t[:, :3], target_class, n_steps, baselines=t[:, 3:], images_per_pass=images_per_pass).flatten(1),
            both, group)
        return flat.view(n, 3, 128, 128)
    out = torch.empty_like(x)
    L = _lib.lib()
    with torch.cuda.device(dev):
        st = _lib.current_stream_ptr()
        for i0 in range(0, n, images_per_pass):
            m = min(images_per_pass, n - i0)
            pts = torch.empty(m * n_steps, 3, 128, 128, dtype=torch.float32, device=dev)
            for j in range(m):
                b = base[i0 + j] if base.shape[0] == n else base[0]
                _lib.check(L.synt_ig_interpolate(x[i0 + j].data_ptr(), b.data_ptr(), n_steps, per,
                                                 pts[j * n_steps:(j + 1) * n_steps].data_ptr(), st), "ig_interpolate")
            _, grads = classifier.score_and_input_gradient(pts, target_class)
            for j in range(m):
                b = base[i0 + j] if base.shape[0] == n else base[0]
                _lib.check(L.synt_ig_reduce(grads[j * n_steps:(j + 1) * n_steps].data_ptr(), x[i0 + j].data_ptr(), b.data_ptr(),
                                            n_steps, per, out[i0 + j].data_ptr(), st), "ig_reduce")
    return out


def compute_combined_attribution(classifier, image, target_class, methods=("ig", "shap"), weights=None, **kwargs):
    """xai/XAI.py:1236-1291: weighted sum of the attribution maps + per-method |attr| mean / max."""
    methods = list(methods)
    if weights is None:
        weights = [1.0 / len(methods)] * len(methods)
    total, details = None, {}
    for method, weight in zip(methods, weights):
        if method == "ig":
            attr = compute_integrated_gradients(classifier, image, target_class, **kwargs.get("ig", {}))
        elif method == "shap":
            attr = compute_shap_approximation(classifier, image, target_class, **kwargs.get("shap", {}))
        elif method == "gradient":
            attr = compute_gradient_attribution(classifier, image, target_class)
        else:
            continue                                             # unknown method: skipped like the reference (XAI.py:1270-1272)
        total = attr * weight if total is None else total + attr * weight
        details[method] = {"weight": weight, "mean_attribution": float(attr.abs().mean()),
                           "max_attribution": float(attr.abs().max())}
    if total is None:
        raise RuntimeError("no attribution could be computed")
    return total, details


# ------------------------------------------------------------------ interventions -------
def counterfactual_intervention_advanced(image, mask, intervention_type="noise", **kwargs):
    """xai/XAI.py:1454-1597: x~ = clamp(x (1-M) + I M, -1, 1).  Extra kwargs: ``noise`` injects the
    N(0,1) tensor, ``generator`` seeds it / the shuffle permutation."""
    noise_std = kwargs.get("noise_std", NOISE_STD)
    blur_kernel = int(kwargs.get("blur_kernel", BLUR_KERNEL_SIZE))
    if blur_kernel % 2 == 0:
        blur_kernel += 1                                        # XAI.py:1512-1513
    dev = image.device
    if not image.is_cuda:
        raise RuntimeError("interventions run on CUDA tensors (no CPU fallback)")
    image = image.contiguous().float()
    B, Cc, H, W = image.shape
    m = torch.from_numpy(mask) if isinstance(mask, np.ndarray) else mask
    m = m.float().to(dev)
    while m.dim() < 3:
        m = m.unsqueeze(0)
    if m.dim() == 4:
        m = m[:, 0]
    m = m.expand(B, H, W).contiguous()
    if intervention_type not in _TYPE_CODES:
        intervention_type = "noise"                             # reference default branch (XAI.py:1563-1565)
    code = _TYPE_CODES[intervention_type]
    aux = None
    gen = kwargs.get("generator")
    if code == 3:
        aux = kwargs.get("noise")
        if aux is None:
            aux = torch.randn(image.shape, device=dev, generator=gen)
        if intervention_type == "gaussian_noise":
            noise_std = max(noise_std, image.std().item() * 0.5)
    elif code == 4:
        # shuffle the masked pixels of every (b, c) plane among themselves (XAI.py:1541-1566), all planes at once: the
        # masked positions in natural order receive the masked values in a random order (sort of uniform keys); the
        # unmasked tail of both index lists only moves pixels that the blend ignores (M = 0 there)
        sel = m.bool().reshape(B, 1, H * W).expand(B, Cc, H * W)
        keys = torch.rand((B, Cc, H * W), device=dev, generator=gen)
        natural = torch.argsort((~sel).to(torch.uint8), dim=-1, stable=True)
        random_ = torch.argsort(torch.where(sel, keys, torch.full_like(keys, 2.0)), dim=-1)
        flat = image.reshape(B, Cc, H * W)
        aux = torch.empty_like(flat).scatter_(-1, natural, flat.gather(-1, random_)).view(B, Cc, H, W)
    out = torch.empty_like(image)
    interv = torch.empty_like(image)
    # 'inpaint' is the reference's fixed 5x5 grouped box convolution (XAI.py:1529-1539); 'blur' honours blur_kernel
    bk = blur_kernel if intervention_type == "blur" else 5
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().synt_intervene_blend_ex(image.data_ptr(), m.data_ptr(),
                                                      aux.contiguous().data_ptr() if aux is not None else None, code,
                                                      float(noise_std), bk, B, Cc, H, W, out.data_ptr(), interv.data_ptr(),
                                                      _lib.current_stream_ptr()), "intervene_blend")
    mask_tensor = m[:, None].expand(B, Cc, H, W)
    diff = (image - out).abs()
    return {
        "modified_image": out,
        "intervention": interv,
        "mask_tensor": mask_tensor,
        "difference": diff,
        "statistics": {
            "intervention_type": intervention_type,
            "mask_coverage": float(mask_tensor.float().mean()),
            "mean_difference": float(diff.mean()),
            "max_difference": float(diff.max()),
            "intervention_strength": float(interv.abs().mean()),
        },
        "parameters": {k: v for k, v in kwargs.items() if k not in ("noise", "generator")},
    }


# ------------------------------------------------------------------ CFI -----------------
def compute_causal_shift_comprehensive(classifier, original_image, modified_image, target_class,
                                       include_all_classes=True, group=None):
    """xai/XAI.py:1600-1700 from ONE batched evaluation of the two images."""
    dev = _dev(classifier)
    pair = torch.cat([original_image.to(dev).reshape(-1, 3, 128, 128)[:1],
                      modified_image.to(dev).reshape(-1, 3, 128, 128)[:1]])
    probs = _probs(classifier, pair, group)
    po, pm = probs[0:1], probs[1:2]
    so, sm = torch.log(po[:, target_class] + 1e-8), torch.log(pm[:, target_class] + 1e-8)
    cfi = so - sm
    delta = cfi.abs() / (so.abs() + 1e-8)
    op, mp = int(po.argmax(1)[0]), int(pm.argmax(1)[0])
    names = CLASS_NAMES
    res = {
        "target_class_analysis": {
            "class_id": target_class, "class_name": names[target_class],
            "cfi": float(cfi), "delta": float(delta),
            "original_score": float(so), "modified_score": float(sm),
            "original_probability": float(po[0, target_class]),
            "modified_probability": float(pm[0, target_class]),
            "probability_shift": float(po[0, target_class] - pm[0, target_class]),
        },
        "prediction_analysis": {
            "original_prediction": op, "original_prediction_name": names[op],
            "modified_prediction": mp, "modified_prediction_name": names[mp],
            "prediction_changed": bool(op != mp),
            "original_confidence": float(po.max()), "modified_confidence": float(pm.max()),
            "confidence_drop": float(po.max() - pm.max()),
        },
    }
    if include_all_classes:
        allc = []
        for cid in range(len(names)):
            a, b = torch.log(po[:, cid] + 1e-8), torch.log(pm[:, cid] + 1e-8)
            allc.append({
                "class_id": cid, "class_name": names[cid], "cfi": float(a - b),
                "delta": float((a - b).abs() / (a.abs() + 1e-8)),
                "original_probability": float(po[0, cid]), "modified_probability": float(pm[0, cid]),
                "probability_shift": float(po[0, cid] - pm[0, cid]),
            })
        res["all_classes_analysis"] = allc
    mid = torch.log((po + pm) / 2 + 1e-8)
    res["distribution_analysis"] = {
        "kl_divergence": float(F.kl_div(torch.log(pm + 1e-8), po, reduction="sum")),
        "js_divergence": float(0.5 * (F.kl_div(mid, po, reduction="sum") + F.kl_div(mid, pm, reduction="sum"))),
        "total_variation": float(0.5 * torch.sum(torch.abs(po - pm))),
    }
    return res


def csi_batch(classifier, images: torch.Tensor, masks: torch.Tensor, intervention_types, target_classes,
              noise: torch.Tensor | None = None, group=None):
    """BASELINE config 5: interventions x ResNet18 inference for a whole batch.  Returns
    cfi[type][b] = s_c(x_b) - s_c(x~_b) with s_c = log(p_c + 1e-8).  With ``group`` the IMAGES are split contiguously over
    the ranks (interventions and evaluations of a slice stay on one rank); one all_gather of the [n, n_types] table."""
    dev = _dev(classifier)
    n_all = images.shape[0]
    rank, world = _world(group)
    lo, hi = shard_bounds(n_all, rank, world) if world > 1 else (0, n_all)
    tc_all = torch.as_tensor(target_classes).long()
    local = None
    if hi > lo:
        imgs = images[lo:hi].to(dev).float().contiguous()
        mk = masks[lo:hi] if masks.dim() == 3 and masks.shape[0] == n_all else masks
        nz = noise[lo:hi] if noise is not None else None
        tc = tc_all[lo:hi].to(dev)
        variants = [imgs]
        for it in intervention_types:
            kw = {"noise": nz.to(dev)} if it in ("noise", "gaussian_noise") and nz is not None else {}
            variants.append(counterfactual_intervention_advanced(imgs, mk.to(dev), it, **kw)["modified_image"])
        probs = _probs(classifier, torch.cat(variants), None).view(len(variants), imgs.shape[0], -1)
        sc = torch.log(probs.gather(2, tc.view(1, -1, 1).expand(len(variants), -1, 1)).squeeze(2) + 1e-8)
        local = (sc[:1] - sc[1:]).t().contiguous()                          # [m, n_types]
    table = gather_rows(local, n_all, group, dev) if world > 1 else local
    return {it: table[:, i] for i, it in enumerate(intervention_types)}


# ------------------------------------------------------------------ statistics ----------
ALPHA_LEVEL = 0.1          # xai/XAI.py:270
N_BOOTSTRAP = 1000         # xai/XAI.py:271
N_PERMUTATIONS = 10000     # xai/XAI.py:272


def draw_bootstrap_indices(n1: int, n2: int, n_bootstrap: int):
    """The resampling indices exactly as the reference's loop draws them from the GLOBAL numpy RNG (XAI.py:1856-1860:
    ``np.random.choice(top_k, n1, replace=True)`` then ``np.random.choice(bottom_k, n2, replace=True)`` per replicate --
    ``choice`` on an array draws ``randint(0, n, size=n)`` indices).  int32 [n_bootstrap, n1], [n_bootstrap, n2]."""
    it = np.empty((n_bootstrap, n1), np.int32)
    ib = np.empty((n_bootstrap, n2), np.int32)
    for b in range(n_bootstrap):
        it[b] = np.random.choice(n1, n1, replace=True)
        ib[b] = np.random.choice(n2, n2, replace=True)
    return it, ib


def draw_permutations(n: int, n_permutations: int) -> np.ndarray:
    """The reference shuffles the SAME pooled array in place again and again (XAI.py:1892-1893), i.e. replicate k sees the
    composition of the first k+1 shuffles: an index array shuffled in place with the same calls reproduces it.  int32 [n_perm, n]."""
    idx = np.arange(n)
    out = np.empty((n_permutations, n), np.int32)
    for k in range(n_permutations):
        np.random.shuffle(idx)
        out[k] = idx
    return out


def resampled_mean_differences(top_k, bottom_k, n_bootstrap=N_BOOTSTRAP, n_permutations=N_PERMUTATIONS, device="cuda",
                               bootstrap_indices=None, permutations=None, seed: int = 0):
    """The two resampling loops of the reference's statistical validation on the GPU (one thread per replicate, float64 means
    in numpy's summation order): returns (bootstrap_diffs [n_bootstrap], permuted_diffs [n_permutations]) as numpy float64.
    ``bootstrap_indices`` = (idx_top, idx_bottom) and ``permutations`` inject the draws (``draw_bootstrap_indices`` /
    ``draw_permutations`` replay the reference's numpy stream); otherwise they come from an in-kernel Philox stream."""
    dev = torch.device(device)
    t = torch.as_tensor(np.asarray(top_k, np.float64), device=dev).contiguous()
    b = torch.as_tensor(np.asarray(bottom_k, np.float64), device=dev).contiguous()
    n1, n2 = t.numel(), b.numel()
    L = _lib.lib()
    boot = torch.empty(n_bootstrap, dtype=torch.float64, device=dev)
    perm = torch.empty(max(n_permutations, 1), dtype=torch.float64, device=dev)
    with torch.cuda.device(dev):
        st = _lib.current_stream_ptr()
        it = ib = None
        if bootstrap_indices is not None:
            it = torch.as_tensor(np.ascontiguousarray(bootstrap_indices[0], np.int32), device=dev)
            ib = torch.as_tensor(np.ascontiguousarray(bootstrap_indices[1], np.int32), device=dev)
        _lib.check(L.synt_stat_bootstrap_mean_diff(t.data_ptr(), n1, b.data_ptr(), n2, it.data_ptr() if it is not None else None,
                                                   ib.data_ptr() if ib is not None else None, int(seed) & 0xFFFFFFFFFFFFFFFF,
                                                   n_bootstrap, boot.data_ptr(), st), "stat_bootstrap")
        if n1 >= 2 and n2 >= 2 and n_permutations > 0:
            comb = torch.cat([t, b]).contiguous()
            pm = torch.as_tensor(np.ascontiguousarray(permutations, np.int32), device=dev) if permutations is not None else None
            _lib.check(L.synt_stat_permutation_mean_diff(comb.data_ptr(), n1 + n2, n1, pm.data_ptr() if pm is not None else None,
                                                         (int(seed) + 1) & 0xFFFFFFFFFFFFFFFF, n_permutations, perm.data_ptr(), st),
                       "stat_permutation")
            permuted = perm[:n_permutations].cpu().numpy()
        else:                                                       # XAI.py:1898-1900: too few values for a permutation test
            permuted = np.array([np.mean(np.asarray(top_k, np.float64)) - np.mean(np.asarray(bottom_k, np.float64))])
    return boot.cpu().numpy(), permuted


def statistical_validation_comprehensive(top_k_shifts, bottom_k_shifts, alpha=ALPHA_LEVEL, n_bootstrap=N_BOOTSTRAP,
                                         n_permutations=N_PERMUTATIONS, device="cuda", bootstrap_indices=None,
                                         permutations=None, seed: int = 0):
    """xai/XAI.py:1708-2005 with the reference's result dictionary.  The bootstrap (1000 resamples) and the permutation test
    (10000 shuffles) -- the Python loops of the reference -- run on the GPU (``resampled_mean_differences``); the closed-form
    tests on the few dozen CFI scalars are the reference's own scipy / numpy calls."""
    from datetime import datetime
    from scipy import stats
    top_k = np.array(top_k_shifts, dtype=np.float64)
    bottom_k = np.array(bottom_k_shifts, dtype=np.float64)

    def describe(data, name):
        return {"name": name, "n": len(data), "mean": np.mean(data), "median": np.median(data), "std": np.std(data, ddof=1),
                "var": np.var(data, ddof=1), "min": np.min(data), "max": np.max(data), "q25": np.percentile(data, 25),
                "q75": np.percentile(data, 75), "iqr": np.percentile(data, 75) - np.percentile(data, 25),
                "skewness": stats.skew(data), "kurtosis": stats.kurtosis(data)}

    descriptive = {"top_k": describe(top_k, "Top-k"), "bottom_k": describe(bottom_k, "Bottom-k")}
    parametric, nonparametric = {}, {}
    t_stat, t_p = stats.ttest_ind(top_k, bottom_k)
    parametric["t_test"] = {"statistic": t_stat, "p_value": t_p, "significant": t_p < alpha, "description": "Independent samples t-test"}
    w_stat, w_p = stats.ttest_ind(top_k, bottom_k, equal_var=False)
    parametric["welch_t_test"] = {"statistic": w_stat, "p_value": w_p, "significant": w_p < alpha,
                                  "description": "Welch's t-test (unequal variances)"}
    u_stat, u_p = stats.mannwhitneyu(top_k, bottom_k, alternative="two-sided")
    nonparametric["mann_whitney_u"] = {"statistic": u_stat, "p_value": u_p, "significant": u_p < alpha, "description": "Mann-Whitney U test"}
    try:
        r_stat, r_p = stats.ranksums(top_k, bottom_k)
        nonparametric["wilcoxon_rank_sum"] = {"statistic": r_stat, "p_value": r_p, "significant": r_p < alpha,
                                              "description": "Wilcoxon rank-sum test"}
    except Exception:                                              # noqa: BLE001  (XAI.py:1806-1807)
        pass
    n1, n2 = len(top_k), len(bottom_k)
    pooled = np.sqrt(((n1 - 1) * np.var(top_k, ddof=1) + (n2 - 1) * np.var(bottom_k, ddof=1)) / (n1 + n2 - 2))
    d = (np.mean(top_k) - np.mean(bottom_k)) / pooled if pooled > 0 else 0
    interp = "negligible" if abs(d) < 0.2 else "small" if abs(d) < 0.5 else "medium" if abs(d) < 0.8 else "large"
    effect = {"cohens_d": {"value": d, "interpretation": interp, "description": "Cohen's d (standardized mean difference)"},
              "glass_delta": {"value": (np.mean(top_k) - np.mean(bottom_k)) / np.std(bottom_k, ddof=1),
                              "description": "Glass's delta (using control group std)"}}
    boot, permuted = resampled_mean_differences(top_k, bottom_k, n_bootstrap, n_permutations, device, bootstrap_indices,
                                                permutations, seed)
    level = 1 - alpha
    ci_lo, ci_hi = np.percentile(boot, (1 - level) / 2 * 100), np.percentile(boot, (1 + level) / 2 * 100)
    bootstrap = {"bootstrap_diffs": boot, "mean_diff": np.mean(boot), "ci_lower": ci_lo, "ci_upper": ci_hi,
                 "ci_contains_zero": ci_lo <= 0 <= ci_hi, "confidence_level": level}
    observed = np.mean(top_k) - np.mean(bottom_k)
    p_perm = np.mean(np.abs(permuted) >= np.abs(observed)) if permuted.size > 1 else 1.0
    permutation = {"observed_difference": observed, "permuted_differences": permuted, "p_value": p_perm,
                   "significant": p_perm < alpha, "n_permutations": n_permutations}
    normality = {}
    try:
        if 3 <= n1 <= 5000 and 3 <= n2 <= 5000:
            st, sb = stats.shapiro(top_k), stats.shapiro(bottom_k)
            normality["shapiro_wilk"] = {"top_k": {"statistic": st[0], "p_value": st[1], "normal": st[1] > alpha},
                                         "bottom_k": {"statistic": sb[0], "p_value": sb[1], "normal": sb[1] > alpha}}
        else:
            normality["shapiro_wilk"] = {"top_k": {"skipped": True, "reason": "sample_size < 3 or > 5000"},
                                         "bottom_k": {"skipped": True, "reason": "sample_size < 3 or > 5000"}}
    except Exception as e:                                         # noqa: BLE001
        normality["shapiro_wilk"] = {"error": str(e)}
    # XAI.py:1934-1935 calls kstest(x, 'norm', args=(mean, std)); scipy >= 1.15 rejects that form for the fast normal CDF, the
    # frozen distribution's cdf is the same test
    kt = stats.kstest(top_k, stats.norm(loc=np.mean(top_k), scale=np.std(top_k)).cdf)
    kb = stats.kstest(bottom_k, stats.norm(loc=np.mean(bottom_k), scale=np.std(bottom_k)).cdf)
    normality["kolmogorov_smirnov"] = {"top_k": {"statistic": kt[0], "p_value": kt[1], "normal": kt[1] > alpha},
                                       "bottom_k": {"statistic": kb[0], "p_value": kb[1], "normal": kb[1] > alpha}}
    lev_stat, lev_p = stats.levene(top_k, bottom_k)
    f_stat = np.var(top_k, ddof=1) / np.var(bottom_k, ddof=1)
    f_p = 2 * min(stats.f.cdf(f_stat, n1 - 1, n2 - 1), 1 - stats.f.cdf(f_stat, n1 - 1, n2 - 1))
    variance = {"levene": {"statistic": lev_stat, "p_value": lev_p, "equal_variances": lev_p > alpha,
                           "description": "Levene's test for equal variances"},
                "f_test": {"statistic": f_stat, "p_value": f_p, "equal_variances": f_p > alpha,
                           "description": "F-test for equal variances"}}
    consensus = {"parametric_significant": any(t["significant"] for t in parametric.values()),
                 "nonparametric_significant": any(t["significant"] for t in nonparametric.values()),
                 "bootstrap_significant": not bootstrap["ci_contains_zero"],
                 "permutation_significant": permutation["significant"]}
    n_sig = sum(consensus.values())
    overall = n_sig >= len(consensus) // 2 + 1
    return {
        "descriptive_statistics": descriptive, "parametric_tests": parametric, "nonparametric_tests": nonparametric,
        "effect_sizes": effect, "bootstrap_analysis": bootstrap, "permutation_analysis": permutation,
        "normality_tests": normality, "variance_tests": variance, "significance_consensus": consensus,
        "overall_conclusion": {"significant": overall, "significant_tests_count": n_sig, "total_tests_count": len(consensus),
                               "alpha_level": alpha, "recommendation": "significant" if overall else "not_significant"},
        "metadata": {"analysis_timestamp": datetime.now().isoformat(), "n_bootstrap_samples": n_bootstrap,
                     "n_permutations": n_permutations, "alpha_level": alpha},
    }


# ------------------------------------------------------------------ analyzer ------------
def select_regions_batch(attributions: torch.Tensor, k_percent=TOP_K_PERCENT, region_type: str = "top",
                         morphology_cleanup: bool = True, connectivity: int = 8):
    """``select_regions_advanced`` for a stack of maps in one launch: ``attributions`` [n,C,H,W] (saliency = channel L2 norm)
    or [n,H,W] (saliency = |x|), H*W <= 16384.  Returns (mask bool [n,H,W], stats float64 [n,8]) on the device; stats
    columns: selected pixels, threshold, mean, std, mean / std / max / min over the selection."""
    if region_type not in ("top", "bottom"):
        raise ValueError(f"unknown region_type {region_type!r}")                    # XAI.py:1384-1385
    if not attributions.is_cuda:
        raise RuntimeError("region selection runs on CUDA tensors (no CPU fallback)")
    a = attributions.detach().float().contiguous()
    use_abs = a.dim() == 3
    n, C = a.shape[0], (1 if use_abs else a.shape[1])
    H, W = a.shape[-2], a.shape[-1]
    mask = torch.empty(n, H, W, dtype=torch.uint8, device=a.device)
    stats = torch.empty(n, 8, dtype=torch.float64, device=a.device)
    with torch.cuda.device(a.device):
        _lib.check(_lib.lib().synt_select_regions(a.data_ptr(), n, C, H, W, int(use_abs), float(k_percent),
                                                  int(region_type == "bottom"), int(bool(morphology_cleanup)), int(connectivity),
                                                  mask.data_ptr(), stats.data_ptr(), _lib.current_stream_ptr()), "select_regions")
    return mask.bool(), stats


def select_regions_advanced(attribution_map, k_percent=TOP_K_PERCENT, region_type="top", morphology_cleanup=True,
                            connectivity=8):
    """xai/XAI.py:1340-1451 with the reference's result dictionary: ``mask`` is a numpy bool array like the reference's
    (its consumers accept it), ``mask_tensor`` the same mask on the device for the GPU consumers.  A 4-D input uses batch
    entry 0, a 3-D input is [C,H,W], a 2-D input is taken by absolute value (XAI.py:1366-1373)."""
    a = attribution_map if torch.is_tensor(attribution_map) else torch.as_tensor(np.asarray(attribution_map))
    shape = tuple(a.shape)
    if not a.is_cuda:
        raise RuntimeError("region selection runs on CUDA tensors (no CPU fallback)")
    if a.dim() == 4:
        a = a[0]
    mask, stats = select_regions_batch(a[None], k_percent, region_type, morphology_cleanup, connectivity)
    st = stats[0].cpu().numpy()
    total = int(a.shape[-1] * a.shape[-2])
    return {
        "mask": mask[0].cpu().numpy(), "mask_tensor": mask[0], "threshold": np.float32(st[1]),
        "statistics": {
            "total_pixels": total, "selected_pixels": int(st[0]), "target_percentage": k_percent,
            "actual_percentage": st[0] / total * 100, "threshold_value": np.float32(st[1]),
            "mean_attribution": st[2], "std_attribution": st[3], "mean_attribution_selected": st[4],
            "std_attribution_selected": st[5], "max_attribution_selected": st[6], "min_attribution_selected": st[7],
        },
        "metadata": {"region_type": region_type, "morphology_cleanup": morphology_cleanup, "connectivity": connectivity,
                     "original_shape": shape},
    }


def _jsonable(o):
    """Result dictionaries for the JSON side-cars: numpy scalars -> Python scalars, replicate arrays -> their length."""
    if isinstance(o, dict):
        return {k: _jsonable(v) for k, v in o.items()}
    if isinstance(o, (list, tuple)):
        return [_jsonable(v) for v in o]
    if isinstance(o, np.ndarray):
        return {"n": int(o.size), "mean": float(o.mean()) if o.size else 0.0}
    if isinstance(o, (np.bool_, bool)):
        return bool(o)
    if isinstance(o, np.integer):
        return int(o)
    if isinstance(o, np.floating):
        return float(o)
    return o


class IntegratedXAIAnalyzer:
    """xai/xai_integration.py:75-132, hot-path stages only: Time-SHAP over every frame, patch-SHAP
    + top/bottom-k interventions + CFI on the key frames [0, T/2, T-4..T-1] (XAI.py:2822-2896)."""

    def __init__(self, device: str = "cuda", verbose: bool = False, precision: str = "bf16", group=None,
                 classifier_path: str | None = None, state_dict: dict | None = None, pretrained: bool = True):
        """``classifier_path`` (a ``torch.save``d state_dict of the classifier or of its ``.model``) / ``state_dict`` load
        trained weights on top; without them the reference's behaviour is kept (ImageNet-pretrained backbone + fresh
        7-way head, xai_integration.py:79) and a warning is raised when the pretrained weights are unavailable offline."""
        self.device = torch.device(device)
        self.verbose = verbose
        self.group = group
        # reference: MelanomaClassifierAdaptive(num_classes=7, architecture='auto', pretrained=True)
        have_weights = classifier_path is not None or state_dict is not None
        self.classifier = MelanomaClassifierAdaptive(num_classes=7, architecture="auto", pretrained=pretrained and not have_weights,
                                                     precision=precision)
        if classifier_path is not None:
            state_dict = torch.load(classifier_path, map_location="cpu")
        if state_dict is not None:
            sd = state_dict.get("model_state_dict", state_dict) if isinstance(state_dict, dict) else state_dict
            if any(k.startswith("model.") for k in sd):
                self.classifier.load_state_dict(sd)
            else:
                self.classifier.model.load_state_dict(sd)
        self.classifier = self.classifier.to(self.device).eval()

    def analyze_trajectory(self, trajectory, class_name, seed, inference_steps, filename, file_path, timesteps=None,
                           shap_samples: int = SHAP_N_SAMPLES, intervention_types=("blur",), methods=("ig", "shap"),
                           ig_steps: int = IG_N_STEPS, stage1_all_frames: bool = True):
        """Stage 1 (XAI.py:2733-2760) runs on EVERY frame like the reference -- IG and patch-SHAP combined 0.5 / 0.5, top /
        bottom 10 % regions -- as batched calls (``compute_integrated_gradients_batch``, ``compute_shap_approximation_batch``,
        ``select_regions_batch``); stage 2 (interventions + CFI, XAI.py:2822-2896) on the key frames [0, T/2, T-4..T-1].
        ``stage1_all_frames=False`` restricts stage 1 to the key frames (what stage 2 consumes)."""
        if not trajectory:
            return None
        T = len(trajectory)
        if timesteps is None or len(timesteps) != T:
            timesteps = list(range(T))                           # xai_integration.py:94-95
        target = CLASS_NAMES.index(class_name) if class_name in CLASS_NAMES else 0
        imp, raw = compute_time_shap(self.classifier, trajectory, timesteps, target, self.group, self.verbose)
        key_frames = sorted(set([0, T // 2] + list(range(max(0, T - 4), T))))
        stage1 = list(range(T)) if stage1_all_frames else key_frames
        frames = torch.cat([trajectory[k].to(self.device).reshape(1, 3, 128, 128).float() for k in stage1])
        methods = [m for m in methods if m in ("ig", "shap", "gradient")]     # unknown methods are skipped (XAI.py:1270-1272)
        if not methods:
            raise RuntimeError("no attribution could be computed")
        attr = None
        for m in methods:                                        # equal weights, XAI.py:2746-2751
            if m == "ig":
                a = compute_integrated_gradients_batch(self.classifier, frames, target, n_steps=ig_steps, group=self.group)
            elif m == "shap":
                a = compute_shap_approximation_batch(self.classifier, frames, target, n_samples=shap_samples, group=self.group)
            else:
                a = compute_gradient_attribution(self.classifier, frames, target)
            attr = a / len(methods) if attr is None else attr + a / len(methods)
        # XAI.py:2753-2760: top / bottom 10 % regions of the combined map, morphology clean-up included, all frames per launch
        top_m, top_s = select_regions_batch(attr, TOP_K_PERCENT, "top")
        bot_m, bot_s = select_regions_batch(attr, BOTTOM_K_PERCENT, "bottom")
        top_s, bot_s = top_s.cpu().numpy(), bot_s.cpu().numpy()
        regions = {f"t_{int(timesteps[k])}": {"top_k": {"selected_pixels": int(top_s[i, 0]), "threshold": float(top_s[i, 1])},
                                              "bottom_k": {"selected_pixels": int(bot_s[i, 0]), "threshold": float(bot_s[i, 1])},
                                              "mean_abs_attribution": float(attr[i].abs().mean())}
                   for i, k in enumerate(stage1)}
        cfi = {}
        for k in key_frames:
            i = stage1.index(k)
            frame = frames[i:i + 1]
            for rname, mask in (("top_k", top_m[i]), ("bottom_k", bot_m[i])):
                for it in intervention_types:
                    mod = counterfactual_intervention_advanced(frame, mask, it)["modified_image"]
                    r = compute_causal_shift_comprehensive(self.classifier, frame, mod, target, group=self.group)
                    cfi[f"t_{k}/{rname}/{it}"] = r["target_class_analysis"]
        # stages 4 / 5 of the reference pipeline (XAI.py:3174-3203): top-k against bottom-k CFI shifts over the key frames
        top_shifts = [v["cfi"] for k, v in cfi.items() if "/top_k/" in k]
        bot_shifts = [v["cfi"] for k, v in cfi.items() if "/bottom_k/" in k]
        statistical = None
        if len(top_shifts) >= 2 and len(bot_shifts) >= 2:
            try:
                statistical = _jsonable(statistical_validation_comprehensive(top_shifts, bot_shifts, device=str(self.device),
                                                                             seed=int(seed) if seed is not None else 0))
            except Exception as e:                                # noqa: BLE001  (the reference logs and goes on, XAI.py:3217-3219)
                statistical = {"error": str(e)}
        return {
            "filename": filename, "file_path": str(file_path), "class_name": class_name, "seed": seed,
            "inference_steps": inference_steps, "n_frames": T,
            "time_shap": {"importance": [float(v) for v in imp],
                          "confidence_scores": [float(v) for v in raw["confidence_scores"]],
                          "probability_scores": [float(v) for v in raw["probability_scores"]],
                          "timesteps": [int(t) for t in timesteps]},
            "region_analysis": regions,
            "cfi": cfi,
            "statistical_validation": statistical,
            "attribution_methods": list(methods),
            "stage1_frames": len(stage1),
            "skipped_stages": ["grad_cam", "plots"],
        }


def create_integrated_xai_analyzer(device: str = "cuda"):
    """xai/xai_integration.py:134"""
    return IntegratedXAIAnalyzer(device=device)


PREVIEW_PATTERNS = ("xai_step_*.png", "gradcam_most_important_*.png", "time_shap_analysis.png")


def run_xai_analysis(image_path, device=None, classifier_path=None, save_dir=None):
    """xai/xai_integration.py:137-159 -- the GUI preview hook: returns ``(PIL RGB image, path)`` of the first stored XAI
    artifact of this image (looked up under ``<two levels above the class folder>/xai_results/<class>/<stem>_*/`` in the
    reference's priority order: step map, Grad-CAM, Time-SHAP plot), else the image itself.  No computation happens here
    (``device`` / ``classifier_path`` / ``save_dir`` are accepted and unused, like in the reference)."""
    from pathlib import Path
    from PIL import Image
    img_path = Path(image_path)
    parents = img_path.parents
    # the reference indexes parents[2] whenever there are >= 2 parents (IndexError for "folder/file.png"); >= 3 here
    root = (parents[2] if len(parents) >= 3 else Path.cwd()) / "xai_results" / img_path.parent.name
    if root.exists():
        for pattern in PREVIEW_PATTERNS:
            hits = sorted(root.glob(f"{img_path.stem}_*/{pattern}"))
            if hits:
                return Image.open(hits[0]).convert("RGB"), str(hits[0])
    return Image.open(img_path).convert("RGB"), str(img_path)
