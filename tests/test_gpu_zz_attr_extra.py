"""Extra GPU checks of the attribution rows; sorted last on purpose.

* Batched Integrated Gradients (pipeline stage 1 of the reference, xai/XAI.py:2740-2751: IG for every frame) against the
  per-image entry point.  fp32 verification mode: the fp32-FMA convolutions compute every output element independently of
  the batch composition, so the two must agree to rounding; bf16: agreement up to the ReLU-mask noise (see
  tests/test_gpu_gradients.py).
* The committed golden fixtures tests/golden/attr.npz (oracle outputs; generator: tests/golden/make_golden_attr.py).
"""
import os
import sys

import numpy as np
import pytest
import torch

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))
import make_golden_attr as mg  # noqa: E402  (seeded inputs of the committed fixtures)

from oracle.classifier import build_classifier
from synt_isic_b200 import MelanomaClassifierAdaptive
from synt_isic_b200 import xai

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("prec", ["fp32", "bf16"])
def test_batched_ig_equals_per_image_ig(cuda_dev, prec):
    oc = build_classifier()
    c = MelanomaClassifierAdaptive(num_classes=7, pretrained=False, precision=prec)
    c.model.load_state_dict(oc.model.state_dict())
    c = c.to(cuda_dev).eval()
    g = torch.Generator().manual_seed(5)
    x = torch.tanh(torch.randn(3, 3, 128, 128, generator=g)).to(cuda_dev)
    base = (torch.randn(1, 3, 128, 128, generator=g) * 0.1).to(cuda_dev)
    got = xai.compute_integrated_gradients_batch(c, x, 1, n_steps=6, baselines=base, images_per_pass=2)
    assert got.shape == x.shape
    for i in range(3):
        one = xai.compute_integrated_gradients(c, x[i:i + 1], 1, n_steps=6, baseline=base)
        a, b = got[i].double().flatten(), one[0].double().flatten()
        rel = ((a - b).norm() / b.norm()).item()
        cosine = (a @ b / (a.norm() * b.norm())).item()
        assert (rel < 1e-3) if prec == "fp32" else (cosine > 0.9), (i, rel, cosine)


def test_golden_attribution_fixtures(cuda_dev):
    """tests/golden/attr.npz (written by tests/golden/make_golden_attr.py from the oracle): region masks / thresholds
    bit-exact, fp32-mode gradient and 20-step Integrated Gradients within the ReLU-flip bound of tests/test_gpu_gradients.py."""
    G = np.load(os.path.join(os.path.dirname(__file__), "golden", "attr.npz"))
    for i, (seed, sigma, kind, conn, k) in enumerate(mg.REGION_CASES):
        got = xai.select_regions_advanced(torch.from_numpy(mg.region_map(seed, sigma)).to(cuda_dev), k, kind, True, conn)
        assert np.array_equal(np.packbits(got["mask"]), G[f"region_mask_{i}"]), i
        assert got["threshold"] == G[f"region_thr_{i}"] and got["statistics"]["selected_pixels"] == int(G[f"region_count_{i}"])
    oc = build_classifier()
    c = MelanomaClassifierAdaptive(num_classes=7, pretrained=False, precision="fp32")
    c.model.load_state_dict(oc.model.state_dict())
    c = c.to(cuda_dev).eval()
    x, base = mg.attribution_inputs()
    grad = xai.compute_gradient_attribution(c, x.to(cuda_dev), 2)[0, :, ::8, ::8].cpu().numpy()
    ig, delta = xai.compute_integrated_gradients(c, x.to(cuda_dev), 2, n_steps=20, baseline=base.to(cuda_dev),
                                                 return_convergence_delta=True)
    ig_sub = ig[0, :, ::8, ::8].cpu().numpy()

    def rel(a, b):
        return float(np.linalg.norm(a.astype(np.float64) - b) / np.linalg.norm(b))

    assert rel(grad, G["grad_sub"]) < 2e-2 and rel(ig_sub, G["ig_sub"]) < 2e-2
    assert abs(float(ig.sum()) - float(G["ig_sum"])) < 5e-3 and abs(delta - float(G["ig_delta"])) < 5e-3


def test_permutation_time_shap_bf16_against_fp32_mode_at_8_steps_8_permutations(cuda_dev):
    """f3 (permutation Time-SHAP over denoising steps) at a size the CPU oracle cannot decode in test time (8 steps, 8
    permutations = 72 coalitions of 8 steps): the production bf16 path against the fp32 verification mode on the same
    injected noise.  The small case against the CPU oracle is tests/test_gpu_classifier.py."""
    from oracle.unet2d import build_unet
    from synt_isic_b200 import DDPMScheduler, SUPPORTED_CONFIG, UNet2DModel
    n, M, target = 8, 8, 1
    oc = build_classifier()
    sd = build_unet(0).state_dict()
    g = torch.Generator().manual_seed(21)
    x_T = torch.randn(1, 3, 128, 128, generator=g).to(cuda_dev)
    noise = torch.randn(n, 1, 3, 128, 128, generator=g).to(cuda_dev)
    out = {}
    for prec in ("fp32", "bf16"):
        model = UNet2DModel(precision=prec, **SUPPORTED_CONFIG)
        model.load_state_dict(sd)
        model = model.to(cuda_dev)
        c = MelanomaClassifierAdaptive(num_classes=7, pretrained=False, precision=prec)
        c.model.load_state_dict(oc.model.state_dict())
        c = c.to(cuda_dev).eval()
        sched = DDPMScheduler(num_train_timesteps=1000, beta_schedule="squaredcos_cap_v2", prediction_type="epsilon")
        sched.set_timesteps(n)
        out[prec] = xai.compute_time_shap_permutation(model, sched, c, x_T, target, n_perm=M, seed=5, noise=noise)
    (phi_f, raw_f), (phi_b, raw_b) = out["fp32"], out["bf16"]
    assert np.array_equal(raw_f["permutations"], raw_b["permutations"]) and raw_f["permutations"].shape == (M, n)
    for phi, raw in (out["fp32"], out["bf16"]):
        assert np.isfinite(phi).all() and abs(raw["efficiency_gap"]) < 1e-9          # sum(phi) = v(all) - v(empty), exactly
        assert np.allclose(raw["prefix_values"][:, 0], raw["prefix_values"][0, 0])   # v(empty) = F(x_T)
    scale = np.abs(raw_f["prefix_values"]).max()
    dv = np.abs(raw_b["prefix_values"] - raw_f["prefix_values"]).max()
    dphi = np.abs(phi_b - phi_f).max()
    print(f"[f3 8x8] |v| max {scale:.4f}  bf16-fp32: prefix values {dv:.3e}  phi {dphi:.3e}  phi range {np.abs(phi_f).max():.3e}")
    # measured on B200: |v| max 0.386, prefix values differ by 3.6e-2 (the ~2e-2 logit error of any bf16 pass), phi by 7.1e-3 of 0.137
    assert dv < 6e-2 and dphi < 2e-2
