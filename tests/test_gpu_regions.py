"""GPU parity of ``select_regions_advanced`` (xai/XAI.py:1340-1451) against the oracle, which
runs the reference's own numpy.percentile / scipy.ndimage calls: masks and thresholds BIT-EXACT, statistics to 1e-5."""
import numpy as np
import pytest
import scipy.ndimage as ndi
import torch

from oracle import xai as oxai
from synt_isic_b200 import xai

pytestmark = pytest.mark.gpu


def make_map(seed, H=128, W=128, sigma=2.0, C=3):
    rng = np.random.default_rng(seed)
    a = rng.standard_normal((C, H, W)).astype(np.float32)
    if sigma:
        a = np.stack([ndi.gaussian_filter(c, sigma) for c in a]).astype(np.float32)
    return a


@pytest.mark.parametrize("sigma", [0, 1.0, 2.5, 4.0])
@pytest.mark.parametrize("kind,conn,k", [("top", 8, 10), ("bottom", 8, 10), ("top", 4, 25), ("bottom", 4, 5)])
def test_masks_and_thresholds_are_bit_exact(cuda_dev, sigma, kind, conn, k):
    a = make_map(int(sigma * 10) + k, sigma=sigma)
    want = oxai.select_regions(a[None], k, kind, True, conn)
    got = xai.select_regions_advanced(torch.from_numpy(a[None]).to(cuda_dev), k, kind, True, conn)
    assert got["threshold"] == want["threshold"]
    assert got["mask"].dtype == np.bool_ and np.array_equal(got["mask"], want["mask"])
    assert torch.equal(got["mask_tensor"].cpu(), torch.from_numpy(want["mask"]))
    for key, w in want["statistics"].items():
        assert abs(float(got["statistics"][key]) - float(w)) <= 1e-5 * max(1.0, abs(float(w))), key
    assert got["metadata"] == want["metadata"]


def test_inputs_shapes_and_batch(cuda_dev):
    # 2-D input = |x|; no clean-up; ragged (non power-of-two) map; a stack of maps in one launch
    a = make_map(3, 64, 96, 1.5)
    for arr, morph in ((a[0], False), (a[0], True), (a, True)):
        want = oxai.select_regions(arr, 10, "top", morph, 8)
        got = xai.select_regions_advanced(torch.from_numpy(arr).to(cuda_dev), 10, "top", morph, 8)
        assert got["threshold"] == want["threshold"] and np.array_equal(got["mask"], want["mask"])
    stack = np.stack([make_map(s, sigma=[0, 2.0, 5.0][s % 3]) for s in range(9)])
    masks, stats = xai.select_regions_batch(torch.from_numpy(stack).to(cuda_dev), 10, "bottom", True, 8)
    assert masks.shape == (9, 128, 128) and stats.shape == (9, 8)
    for i in range(9):
        want = oxai.select_regions(stack[i], 10, "bottom", True, 8)
        assert np.array_equal(masks[i].cpu().numpy(), want["mask"]), i
        assert stats[i, 1].item() == float(want["threshold"])
    # patch-constant map (what patch-SHAP produces): massive ties at the threshold
    blocks = np.repeat(np.repeat(np.random.default_rng(1).standard_normal((1, 8, 8)).astype(np.float32), 16, 1), 16, 2)
    blocks = np.repeat(blocks, 3, 0)
    want = oxai.select_regions(blocks, 10, "top", True, 8)
    got = xai.select_regions_advanced(torch.from_numpy(blocks).to(cuda_dev), 10, "top", True, 8)
    assert got["threshold"] == want["threshold"] and np.array_equal(got["mask"], want["mask"])
    # everything-selected / nothing-survives corner cases
    flat = np.ones((3, 128, 128), np.float32)
    want = oxai.select_regions(flat, 10, "top", True, 8)
    got = xai.select_regions_advanced(torch.from_numpy(flat).to(cuda_dev), 10, "top", True, 8)
    assert np.array_equal(got["mask"], want["mask"])
    noise = make_map(11, sigma=0)
    want = oxai.select_regions(noise, 5, "top", True, 8)
    got = xai.select_regions_advanced(torch.from_numpy(noise).to(cuda_dev), 5, "top", True, 8)
    assert np.array_equal(got["mask"], want["mask"]) and got["statistics"]["selected_pixels"] == want["statistics"]["selected_pixels"]
    with pytest.raises(ValueError):
        xai.select_regions_advanced(torch.from_numpy(noise).to(cuda_dev), 5, "middle")
    with pytest.raises(RuntimeError):
        xai.select_regions_advanced(torch.from_numpy(noise), 5, "top")
