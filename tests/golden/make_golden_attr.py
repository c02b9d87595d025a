"""Generates tests/golden/attr.npz from the ORACLE for the attribution rows:

    python tests/golden/make_golden_attr.py

  grad_sub / ig_sub   d log(p_c + 1e-8)/dx and the 20-step riemann_right Integrated Gradients map of the oracle classifier
                      (seed-7 weights, torch autograd) for a seeded image and baseline, sub-sampled [::8, ::8]; ig_delta
  region_*            select_regions (the reference's numpy.percentile / scipy.ndimage calls) on five seeded maps:
                      bit-packed masks, float32 thresholds, selected-pixel counts
"""
import os
import sys

import numpy as np
import scipy.ndimage as ndi
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import xai as oxai  # noqa: E402
from oracle.classifier import build_classifier  # noqa: E402

REGION_CASES = [(0, 5.0, "top", 8, 10), (1, 1.0, "bottom", 8, 10), (2, 3.0, "top", 4, 25), (3, 6.0, "bottom", 4, 5),
                (4, 0.0, "top", 8, 5)]                      # (seed, gaussian sigma, region_type, connectivity, k_percent); last: nothing survives


def region_map(seed, sigma):
    a = np.random.default_rng(1000 + seed).standard_normal((3, 128, 128)).astype(np.float32)
    return np.stack([ndi.gaussian_filter(c, sigma) for c in a]).astype(np.float32) if sigma else a


def attribution_inputs():
    g = torch.Generator().manual_seed(4242)
    x = torch.tanh(torch.randn(1, 3, 128, 128, generator=g))
    base = torch.randn(1, 3, 128, 128, generator=g) * 0.1
    return x, base


def main():
    torch.set_num_threads(os.cpu_count())
    c = build_classifier()
    x, base = attribution_inputs()
    grad = oxai.gradient_attribution(c, x, 2)
    ig, delta = oxai.integrated_gradients(c, x, 2, base, n_steps=20)
    out = {"grad_sub": grad[0, :, ::8, ::8].numpy(), "ig_sub": ig[0, :, ::8, ::8].numpy(), "ig_delta": np.array(delta),
           "ig_sum": np.array(float(ig.sum()))}
    for i, (seed, sigma, kind, conn, k) in enumerate(REGION_CASES):
        r = oxai.select_regions(region_map(seed, sigma), k, kind, True, conn)
        out[f"region_mask_{i}"] = np.packbits(r["mask"])
        out[f"region_thr_{i}"] = np.array(r["threshold"], dtype=np.float32)
        out[f"region_count_{i}"] = np.array(r["statistics"]["selected_pixels"])
    np.savez(os.path.join(HERE, "attr.npz"), **out)
    print("written", os.path.join(HERE, "attr.npz"), {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
