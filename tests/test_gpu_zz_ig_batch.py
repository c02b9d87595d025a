"""Batched Integrated Gradients (pipeline stage 1 of the reference, xai/XAI.py:2740-2751: IG for every frame) against the
per-image entry point.  fp32 verification mode: the fp32-FMA convolutions compute every output element independently of
the batch composition, so the two must agree to rounding; bf16: agreement up to the ReLU-mask noise (see
tests/test_gpu_gradients.py)."""
import pytest
import torch

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
