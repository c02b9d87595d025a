"""CPU: the oracle against the committed golden fixtures and the indirect reference pins of
SURVEY.md section 8(c) (the reference has no tests; diffusers is not installable here ->
"parity unpinned" for the UNet/scheduler restatement; these are the pins that exist)."""
import hashlib
import json
import os

import numpy as np
import pytest
import torch

from oracle import xai as oxai
from oracle.classifier import CLASS_NAMES, build_classifier
from oracle.ddpm import DDPMSchedulerOracle
from oracle.unet2d import EXPECTED_PARAM_COUNT, build_unet

G = os.path.join(os.path.dirname(__file__), "golden")


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


@pytest.fixture(scope="module")
def unet():
    return build_unet(0)


def test_param_count_matches_checkpoint_size_pin(unet):
    # core/cache/metadata/cache_metadata.json:7 -> 101,345,019 B = 25,304,963 fp32 + pickle overhead
    n = sum(p.numel() for p in unet.parameters())
    assert n == EXPECTED_PARAM_COUNT == 25_304_963
    assert 101_345_019 - 4 * n < 200_000


def test_state_dict_keys_follow_diffusers_scheme(unet):
    keys = set(unet.state_dict())
    assert len(keys) == 330
    for k in ["conv_in.weight", "time_embedding.linear_2.bias", "down_blocks.0.resnets.1.time_emb_proj.weight",
              "down_blocks.1.resnets.0.conv_shortcut.weight", "down_blocks.2.attentions.1.to_out.0.bias",
              "down_blocks.2.downsamplers.0.conv.weight", "mid_block.attentions.0.group_norm.weight",
              "up_blocks.1.attentions.2.to_q.weight", "up_blocks.2.upsamplers.0.conv.bias",
              "up_blocks.3.resnets.2.conv_shortcut.bias", "conv_norm_out.weight", "conv_out.bias"]:
        assert k in keys, k
    assert "down_blocks.3.downsamplers.0.conv.weight" not in keys
    assert "up_blocks.3.upsamplers.0.conv.weight" not in keys
    assert unet.state_dict()["up_blocks.2.resnets.2.conv1.weight"].shape == (128, 192, 3, 3)
    assert unet.state_dict()["up_blocks.1.resnets.2.conv1.weight"].shape == (256, 384, 3, 3)


def test_unet_shape_smoke_like_reference(unet):
    # xai/XAI.py:608-615: randn(1,3,128,128), randint(0,1000,(1,)) -> same shape
    x = torch.randn(1, 3, 128, 128)
    with torch.no_grad():
        y = unet(x, torch.randint(0, 1000, (1,))).sample
    assert y.shape == x.shape and torch.isfinite(y).all()


def test_unet_golden(unet):
    gold = np.load(os.path.join(G, "unet_eps.npz"))
    g = torch.Generator().manual_seed(2024)
    x = torch.randn(2, 3, 128, 128, generator=g)
    with torch.no_grad():
        for t in (980, 0):
            got = unet(x, t).sample[:, :, ::16, ::16].numpy()
            np.testing.assert_allclose(got, gold[f"eps_t{t}"], rtol=0, atol=2e-5)


def test_scheduler_known_answers_and_golden():
    tab = json.load(open(os.path.join(G, "ddpm_tables.json")))
    s = DDPMSchedulerOracle()
    # SURVEY.md A.3 known answers
    assert abs(float(s.betas[0]) - 4.12842237e-05) < 1e-12
    assert abs(float(s.betas[998]) - 0.749999404) < 1e-8
    assert float(s.betas[999]) == pytest.approx(0.999, abs=1e-7)
    assert abs(float(s.alphas_cumprod[500]) - 0.492285043) < 1e-8
    assert abs(float(s.alphas_cumprod[999]) - 2.428734991e-09) < 1e-15
    assert sha(s.alphas_cumprod.numpy()) == tab["acp_sha256"]
    assert sha(s.betas.numpy()) == tab["betas_sha256"]
    s.set_timesteps(50)
    assert s.timesteps.tolist() == list(range(980, -1, -20)) == tab["timesteps_50"]     # XAI.py:742-748: high -> 0
    assert abs(float(s.coefficients(980)[4] ** 2) - 0.76028657) < 1e-7
    assert abs(float(s.coefficients(20)[4] ** 2) - 4.0402585e-05) < 1e-11
    s.set_timesteps(1000)
    assert s.timesteps.tolist() == list(range(999, -1, -1))
    assert abs(float(s.coefficients(999)[4] ** 2) - 0.99899757) < 1e-7
    s.set_timesteps(7)
    assert s.timesteps.tolist() == [852, 710, 568, 426, 284, 142, 0] == tab["timesteps_7"]
    for n in (50, 1000, 7):
        s.set_timesteps(n)
        coef = np.array([[float(v) for v in s.coefficients(int(t))] for t in s.timesteps], dtype=np.float32)
        coef[-1, 4] = 0.0
        assert sha(coef) == tab[f"coef_sha256_{n}"]


def test_scheduler_step_semantics():
    s = DDPMSchedulerOracle()
    s.set_timesteps(50)
    g = torch.Generator().manual_seed(0)
    x, eps, z = (torch.randn(1, 3, 8, 8, generator=g) for _ in range(3))
    out = s.step(eps, 980, x, noise=z).prev_sample
    sb, sa, c0, ct, sig = s.coefficients(980)
    x0 = ((x - sb * eps) / sa).clamp(-1, 1)
    assert torch.equal(out, c0 * x0 + ct * x + sig * z)
    last = s.step(eps, 0, x, noise=z)                      # t == 0: no noise
    assert torch.equal(last.prev_sample, s.coefficients(0)[2] * last.pred_original_sample + s.coefficients(0)[3] * x)


def test_md5_seed_offsets():
    tab = json.load(open(os.path.join(G, "ddpm_tables.json")))
    want = {"MEL": 2133561680, "NV": 1396962837, "BCC": 533065696, "AKIEC": 189026585, "BKL": 438814178,
            "DF": 965706499, "VASC": 1149163796}
    assert tab["md5_offsets"] == want


def test_classifier_golden_and_smoke():
    gold = np.load(os.path.join(G, "classifier.npz"))
    c = build_classifier()
    g = torch.Generator().manual_seed(2024)
    torch.randn(2, 3, 128, 128, generator=g)                       # same stream position as make_golden
    imgs = torch.tanh(torch.randn(4, 3, 128, 128, generator=g) * 1.5)
    with torch.no_grad():
        logits = c(imgs)
        probs = c.get_probabilities(imgs)
    assert logits.shape == (4, 7)                                  # XAI.py:546-551
    assert torch.allclose(probs.sum(1), torch.ones(4), atol=1e-5)
    np.testing.assert_allclose(logits.numpy(), gold["logits"], atol=1e-4)
    assert CLASS_NAMES == ["MEL", "NV", "BCC", "AKIEC", "BKL", "DF", "VASC"]


def test_antialias_is_noop_when_upsampling():
    import torch.nn.functional as F
    x = torch.rand(1, 3, 128, 128)
    a = F.interpolate(x, size=(224, 224), mode="bilinear", align_corners=False, antialias=True)
    b = F.interpolate(x, size=(224, 224), mode="bilinear", align_corners=False, antialias=False)
    assert (a - b).abs().max() < 1e-5


def test_xai_oracle_golden():
    gold = np.load(os.path.join(G, "xai.npz"))
    c = build_classifier()
    g = torch.Generator().manual_seed(2024)
    torch.randn(2, 3, 128, 128, generator=g)
    torch.randn(4, 3, 128, 128, generator=g)
    traj = [torch.tanh(torch.randn(1, 3, 128, 128, generator=g)) for _ in range(6)]
    imp, raw = oxai.time_shap(c, traj, list(range(6)), 0)
    np.testing.assert_allclose(imp, gold["time_shap"], atol=1e-4)
    assert imp.min() == 0.0 and imp.max() == 1.0 and len(imp) == 6           # XAI.py:1304-1333
    masks = torch.from_numpy(gold["patch_masks"])
    attr = oxai.patch_shap(c, traj[-1], 0, masks)
    assert attr.shape == traj[-1].shape
    np.testing.assert_allclose(attr[0, 0, ::16, ::16].numpy(), gold["patch_attr_sub"], atol=1e-5)
    mask = torch.zeros(128, 128)
    mask[40:80, 30:90] = 1
    blur, _ = oxai.intervene(traj[-1], mask, "blur")
    np.testing.assert_allclose(blur[0, :, ::8, ::8].numpy(), gold["blur_sub"], atol=1e-6)
    cs = oxai.causal_shift(c, traj[-1], blur, 0)
    np.testing.assert_allclose([cs["cfi"], cs["delta"], cs["kl_divergence"]], gold["cfi"], atol=1e-4)


def test_permutation_shapley_estimator_and_enumeration():
    """Permutation Time-SHAP (README.md:171-221 of the reference, spec only): the coalition enumeration is a pure function
    of (n, M, seed) and identical in the oracle and in the product; with ALL n! permutations the estimator equals the
    brute-force Shapley values of a set function with interactions; efficiency holds for any M."""
    import itertools
    from oracle import xai as oxai
    from synt_isic_b200 import xai
    p1, p2 = oxai.step_permutations(7, 5, 123), xai.draw_step_permutations(7, 5, 123)
    assert p1.dtype == np.int64 and np.array_equal(p1, p2)                       # bit-exact enumeration
    assert np.array_equal(p1[0], np.random.default_rng(123).permutation(7))
    assert all(sorted(r) == list(range(7)) for r in p1.tolist())
    m = xai.prefix_coalitions(p1[0])
    assert m.shape == (8, 7) and m[0].sum() == 0 and m[7].sum() == 7 and all(m[k + 1, p1[0][k]] == 1 for k in range(7))
    w = np.array([0.5, -1.0, 2.0, 0.25])
    def v(S):                                                                    # additive part + one pairwise interaction
        return float(sum(w[i] for i in S) + (3.0 if {1, 2} <= set(S) else 0.0))
    exact = oxai.exact_shapley(v, 4)
    assert np.allclose(exact, [0.5, 0.5, 3.5, 0.25])
    perms = np.array(list(itertools.permutations(range(4))), np.int64)
    values = np.array([[v(set(p[:k])) for k in range(5)] for p in perms])
    assert np.allclose(xai.shapley_from_prefix_values(perms, values), exact)
    few = xai.draw_step_permutations(4, 3, 9)
    vals = np.array([[v(set(p[:k])) for k in range(5)] for p in few])
    assert abs(xai.shapley_from_prefix_values(few, vals).sum() - (v({0, 1, 2, 3}) - v(set()))) < 1e-12


def test_oracle_integrated_gradients_completeness_and_gradient_taps():
    """captum riemann_right restatement (oracle/xai.py): completeness axiom sum(IG) ~= F(x) - F(x') up to the Riemann error,
    IG of x' == x is zero, and the tap helper's input gradient equals plain autograd."""
    from oracle import xai as oxai
    from oracle.classifier import build_classifier
    m = build_classifier()
    g = torch.Generator().manual_seed(3)
    x = torch.tanh(torch.randn(1, 3, 128, 128, generator=g))
    base = torch.randn(1, 3, 128, 128, generator=g) * 0.1
    attr, delta = oxai.integrated_gradients(m, x, 2, base, n_steps=20)
    with torch.no_grad():
        diff = float(m.get_per_class_score(x, 2) - m.get_per_class_score(base, 2))
    assert abs(delta) < 0.15 * abs(diff) + 0.02
    assert abs(float(attr.sum()) - diff - delta) < 1e-4
    zero, _ = oxai.integrated_gradients(m, x, 2, x.clone(), n_steps=4)
    assert float(zero.abs().max()) == 0.0
    taps, score = oxai.classifier_gradient_taps(m, x, 2)
    assert torch.allclose(taps["input"][1], oxai.gradient_attribution(m, x, 2), atol=1e-7)
    assert set(taps) >= {"preprocess", "relu", "maxpool", "layer1.0", "layer4.1", "input"}


def test_attribution_golden_fixtures():
    """tests/golden/attr.npz pins the attribution oracles: gradient / Integrated Gradients of the torchvision classifier and
    the numpy-percentile / scipy-morphology region selection (a numpy or scipy upgrade that changes a mask shows up here)."""
    import sys
    sys.path.insert(0, G)
    import make_golden_attr as mg
    from oracle import xai as oxai
    from oracle.classifier import build_classifier
    fx = np.load(os.path.join(G, "attr.npz"))
    for i, (seed, sigma, kind, conn, k) in enumerate(mg.REGION_CASES):
        r = oxai.select_regions(mg.region_map(seed, sigma), k, kind, True, conn)
        assert np.array_equal(np.packbits(r["mask"]), fx[f"region_mask_{i}"]), i
        assert r["threshold"] == fx[f"region_thr_{i}"] and r["statistics"]["selected_pixels"] == int(fx[f"region_count_{i}"])
    m = build_classifier()
    x, base = mg.attribution_inputs()
    grad = oxai.gradient_attribution(m, x, 2)[0, :, ::8, ::8].numpy()
    assert np.abs(grad - fx["grad_sub"]).max() <= 1e-4 * np.abs(fx["grad_sub"]).max()
    ig, delta = oxai.integrated_gradients(m, x, 2, base, n_steps=20)
    assert np.abs(ig[0, :, ::8, ::8].numpy() - fx["ig_sub"]).max() <= 1e-4 * np.abs(fx["ig_sub"]).max()
    assert abs(delta - float(fx["ig_delta"])) < 1e-4
