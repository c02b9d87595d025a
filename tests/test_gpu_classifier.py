"""GPU parity of the ResNet18 logit path and the XAI evaluation loops against the oracle
(real torchvision resnet18 + the reference's preprocess), through the drop-in objects / C ABI."""
import numpy as np
import pytest
import torch

from oracle import xai as oxai
from oracle.classifier import build_classifier
from synt_isic_b200 import MelanomaClassifierAdaptive
from synt_isic_b200 import xai

pytestmark = pytest.mark.gpu


def rel(a, b):
    return ((a.double() - b.double()).norm() / b.double().norm()).item()


@pytest.fixture(scope="module")
def oc():
    return build_classifier()


@pytest.fixture(scope="module")
def clfs(oc, cuda_dev):
    out = {}
    for prec in ("fp32", "bf16"):
        c = MelanomaClassifierAdaptive(num_classes=7, pretrained=False, precision=prec)
        c.model.load_state_dict(oc.model.state_dict())
        out[prec] = c.to(cuda_dev).eval()
    return out


@pytest.fixture(scope="module")
def frames():
    g = torch.Generator().manual_seed(77)
    return torch.tanh(torch.randn(12, 3, 128, 128, generator=g) * 1.5)


@pytest.mark.parametrize("prec,tol", [("fp32", 2e-5), ("bf16", 2e-2)])
def test_logits(oc, clfs, frames, cuda_dev, prec, tol):
    with torch.no_grad():
        ref = oc(frames)
    got = clfs[prec](frames.to(cuda_dev)).cpu()
    assert got.shape == (12, 7)
    assert rel(got, ref) <= tol
    probs = clfs[prec].get_probabilities(frames.to(cuda_dev))
    assert torch.allclose(probs.sum(1).cpu(), torch.ones(12), atol=1e-5)          # XAI.py:546-551
    one = clfs[prec](frames[5:6].to(cuda_dev)).cpu()                              # odd batch vs batch of 12
    assert rel(one, got[5:6]) < (1e-5 if prec == "fp32" else 5e-3)


def test_preprocess_kernel(oc, clfs, frames, cuda_dev):
    ref = oc.preprocess_for_classifier(frames[:2] * 1.3)                          # exercises the clamp
    got = clfs["fp32"].debug_tap((frames[:2] * 1.3).to(cuda_dev), "preprocess").cpu()
    assert (got - ref).abs().max().item() < 2e-5


def test_weight_mutation_invalidates_folded_weights(oc, cuda_dev, frames):
    c = MelanomaClassifierAdaptive(num_classes=7, pretrained=False, precision="fp32")
    c.model.load_state_dict(oc.model.state_dict())
    c = c.to(cuda_dev).eval()
    a = c(frames[:2].to(cuda_dev))
    with torch.no_grad():
        c.model.fc.weight.mul_(0.5)                                               # XAI.py:2055-2059
    b = c(frames[:2].to(cuda_dev))
    assert not torch.equal(a, b)


@pytest.mark.parametrize("prec,tol", [("fp32", 1e-4), ("bf16", 1e-3)])
def test_time_shap(oc, clfs, frames, cuda_dev, prec, tol):
    """Time-SHAP values within 1e-3 absolute (north_star) -- config 4."""
    traj = [frames[i:i + 1] for i in range(12)]
    ref_imp, ref_raw = oxai.time_shap(oc, traj, list(range(12)), 2)
    imp, raw = xai.compute_time_shap(clfs[prec], [f.to(cuda_dev) for f in traj], list(range(12)), 2)
    assert imp.shape == (12,) and imp.dtype == np.float64
    # scores are log-probabilities: compare them absolutely, and the min-max normalised importance
    np.testing.assert_allclose(raw["confidence_scores"], ref_raw["confidence_scores"], atol=tol * 20)
    np.testing.assert_allclose(raw["probability_scores"], ref_raw["probability_scores"], atol=tol)
    if prec == "fp32":
        np.testing.assert_allclose(imp, ref_imp, atol=1e-3)
    assert imp.min() == 0.0 and imp.max() == 1.0


def test_patch_shap_coalitions(oc, clfs, frames, cuda_dev):
    g = torch.Generator().manual_seed(1)
    masks = torch.rand(24, 8, 8, generator=g) > 0.5
    ref = oxai.patch_shap(oc, frames[:1], 1, masks)
    got = xai.compute_shap_approximation(clfs["fp32"], frames[:1].to(cuda_dev), 1, patch_masks=masks).cpu()
    assert got.shape == ref.shape
    assert (got - ref).abs().max().item() < 1e-4
    # coalition enumeration (mask expansion) is bit-exact
    pm = masks.to(cuda_dev).to(torch.uint8)
    from synt_isic_b200 import _lib
    out = torch.empty(24, 3, 128, 128, device=cuda_dev)
    x0 = frames[0].to(cuda_dev).contiguous()
    _lib.check(_lib.lib().synt_patch_mask_apply(x0.data_ptr(), pm.data_ptr(), 24, 3, 128, 128, 16, out.data_ptr(),
                                                _lib.current_stream_ptr()))
    for i in range(24):
        full = oxai.expand_patch_mask(masks[i])
        want = frames[0].clone()
        want[:, ~full] = 0
        assert torch.equal(out[i].cpu(), want)
    # masks drawn like the reference (global CPU RNG stream)
    torch.manual_seed(3)
    a = xai.draw_patch_masks(5)
    torch.manual_seed(3)
    b = torch.stack([torch.rand(8, 8) > 0.5 for _ in range(5)])
    assert torch.equal(a, b)


@pytest.mark.parametrize("kind", ["zero", "mean", "blur", "noise", "inpaint"])
def test_interventions(frames, cuda_dev, kind):
    g = torch.Generator().manual_seed(2)
    mask = (torch.rand(128, 128, generator=g) > 0.8)
    noise = torch.randn(2, 3, 128, 128, generator=g)
    img = frames[:2]
    ref, _ = oxai.intervene(img, mask, kind, noise=noise)
    got = xai.counterfactual_intervention_advanced(img.to(cuda_dev), mask.numpy(), kind, noise=noise.to(cuda_dev))
    assert (got["modified_image"].cpu() - ref).abs().max().item() < 2e-6
    assert got["statistics"]["intervention_type"] == kind
    assert abs(got["statistics"]["mask_coverage"] - mask.float().mean().item()) < 1e-6


def test_shuffle_intervention_permutes_masked_pixels(frames, cuda_dev):
    mask = torch.zeros(128, 128); mask[10:50, 20:70] = 1
    img = frames[:1].to(cuda_dev)
    out = xai.counterfactual_intervention_advanced(img, mask, "shuffle")["modified_image"]
    sel = mask.bool().to(cuda_dev)
    assert torch.equal(out[0][:, ~sel], img[0][:, ~sel])
    for c in range(3):
        assert torch.equal(out[0, c][sel].sort().values, img[0, c][sel].sort().values)
        assert not torch.equal(out[0, c][sel], img[0, c][sel])


def test_causal_shift(oc, clfs, frames, cuda_dev):
    mask = torch.zeros(128, 128); mask[40:80, 30:90] = 1
    blur, _ = oxai.intervene(frames[3:4], mask, "blur")
    ref = oxai.causal_shift(oc, frames[3:4], blur, 0)
    got = xai.compute_causal_shift_comprehensive(clfs["fp32"], frames[3:4].to(cuda_dev), blur.to(cuda_dev), 0)
    t = got["target_class_analysis"]
    assert abs(t["cfi"] - ref["cfi"]) < 1e-4 and abs(t["delta"] - ref["delta"]) < 1e-4
    assert got["prediction_analysis"]["prediction_changed"] == ref["prediction_changed"]
    assert abs(got["distribution_analysis"]["kl_divergence"] - ref["kl_divergence"]) < 1e-5
    assert len(got["all_classes_analysis"]) == 7
    np.testing.assert_allclose([c["cfi"] for c in got["all_classes_analysis"]], ref["all_cfi"], atol=1e-4)


def test_csi_batch_config5(oc, clfs, cuda_dev):
    """BASELINE configs[4]: interventions x ResNet18 inference on a batch (64 here, 256 in bench)."""
    g = torch.Generator().manual_seed(5)
    imgs = torch.tanh(torch.randn(64, 3, 128, 128, generator=g))
    masks = (torch.rand(64, 128, 128, generator=g) > 0.9).float()
    noise = torch.randn(64, 3, 128, 128, generator=g)
    tc = [i % 7 for i in range(64)]
    got = xai.csi_batch(clfs["fp32"], imgs.to(cuda_dev), masks.to(cuda_dev), ["blur", "zero", "mean", "noise"], tc,
                        noise=noise.to(cuda_dev))
    for kind in ("blur", "noise"):
        for b in (0, 17, 63):
            mod, _ = oxai.intervene(imgs[b:b + 1], masks[b], kind, noise=noise[b:b + 1])
            r = oxai.causal_shift(oc, imgs[b:b + 1], mod, tc[b])
            assert abs(got[kind][b].item() - r["cfi"]) < 2e-4


def test_integrated_analyzer_runs(cuda_dev, frames):
    an = xai.IntegratedXAIAnalyzer(device="cuda:0", precision="bf16")
    traj = [frames[i:i + 1].to(cuda_dev) for i in range(6)]
    res = an.analyze_trajectory(traj, "NV", 42, 6, "ISIC_0000001.png", "/tmp/ISIC_0000001.png", shap_samples=16)
    assert res["n_frames"] == 6 and len(res["time_shap"]["importance"]) == 6
    assert any(k.startswith("t_5/top_k/blur") for k in res["cfi"])
    import json
    json.dumps(res)


def test_fused_front_end_is_bit_identical_to_the_three_kernels(oc, cuda_dev, frames, monkeypatch):
    """bf16 mode: preprocess + 7x7/s2 stem + maxpool in ONE kernel (tcgen05 by default, mma.sync with SYNT_STEM_TC=0) against the unfused chain
    (SYNT_RESNET_FUSE_FRONT=0) and against the im2col + tcgen05 stem (SYNT_STEM_IM2COL=1): same arithmetic per output
    element, hence identical logits for the first two and bf16-level agreement with the third
    (classifier of xai/XAI.py:357-471)."""
    def build(env):
        for k, v in env.items():
            monkeypatch.setenv(k, v)
        c = MelanomaClassifierAdaptive(num_classes=7, pretrained=False, precision="bf16")
        c.model.load_state_dict(oc.model.state_dict())
        c = c.to(cuda_dev).eval()
        y = c(frames.to(cuda_dev)).cpu()                      # creates the native handle under this environment
        for k in env:
            monkeypatch.delenv(k)
        return y
    tc = build({})                                            # default: the tcgen05 front end (space-to-depth taps)
    fused = build({"SYNT_STEM_TC": "0"})                      # the mma.sync fused front end
    chain = build({"SYNT_RESNET_FUSE_FRONT": "0"})
    im2col = build({"SYNT_RESNET_FUSE_FRONT": "0", "SYNT_STEM_IM2COL": "1"})
    assert torch.equal(fused, chain)
    assert rel(im2col, fused) < 5e-3
    assert rel(tc, fused) < 5e-3                              # same bf16 operands, another fp32 summation order


def test_time_shap_streams_a_host_trajectory(clfs, cuda_dev):
    """A host-resident (pinned) trajectory of more than 256 frames is streamed through the classifier in chunks with the
    copies on a side stream: same importances and raw scores as the on-device call (xai/XAI.py:1179-1234)."""
    g = torch.Generator().manual_seed(5)
    traj = torch.tanh(torch.randn(300, 3, 128, 128, generator=g)).pin_memory()
    imp_h, raw_h = xai.compute_time_shap(clfs["bf16"], traj, list(range(300)), 3)
    imp_d, raw_d = xai.compute_time_shap(clfs["bf16"], traj.to(cuda_dev), list(range(300)), 3)
    assert np.allclose(raw_h["confidence_scores"], raw_d["confidence_scores"], atol=1e-6)
    assert np.allclose(imp_h, imp_d, atol=1e-6)


def test_permutation_time_shap_matches_the_oracle(oc, clfs, cuda_dev):
    """Permutation Shapley over denoising steps (README.md:171-221 of the reference): per permutation the n+1 prefix
    coalitions are decoded as ONE batch with a per-image step mask in the fused scheduler epilogue and scored in one
    classifier call; against the oracle's one-decode-per-coalition loops on the same weights, x_T and noise (fp32 mode).
    Also: efficiency (sum phi = v(all) - v(empty)) and the in-kernel shared noise field (bf16 path runs)."""
    from oracle.ddpm import DDPMSchedulerOracle
    from oracle.unet2d import build_unet
    from synt_isic_b200 import DDPMScheduler, SUPPORTED_CONFIG, UNet2DModel
    n, M, target = 3, 2, 1
    ounet = build_unet(0)
    osched = DDPMSchedulerOracle(); osched.set_timesteps(n)
    g = torch.Generator().manual_seed(11)
    x_T = torch.randn(1, 3, 128, 128, generator=g)
    noise = torch.randn(n, 1, 3, 128, 128, generator=g)
    phi_ref, val_ref = oxai.time_shap_permutation(ounet, osched, oc, x_T, target, M, 5, noise)
    model = UNet2DModel(precision="fp32", **SUPPORTED_CONFIG)
    model.load_state_dict(ounet.state_dict())
    model = model.to(cuda_dev)
    sched = DDPMScheduler(num_train_timesteps=1000, beta_schedule="squaredcos_cap_v2", prediction_type="epsilon")
    sched.set_timesteps(n)
    phi, raw = xai.compute_time_shap_permutation(model, sched, clfs["fp32"], x_T.to(cuda_dev), target, n_perm=M, seed=5,
                                                 noise=noise.to(cuda_dev))
    assert np.array_equal(raw["permutations"], oxai.step_permutations(n, M, 5))
    assert np.allclose(raw["prefix_values"], val_ref, atol=2e-3), np.abs(raw["prefix_values"] - val_ref).max()
    assert np.allclose(phi, phi_ref, atol=2e-3)
    assert abs(raw["efficiency_gap"]) < 1e-9
    assert np.allclose(raw["prefix_values"][:, 0], raw["prefix_values"][0, 0])           # v(empty) = F(x_T), no decode
    # production path: bf16 tensor kernels, in-kernel Philox noise shared by all coalitions
    mb = UNet2DModel(precision="bf16", **SUPPORTED_CONFIG)
    mb.load_state_dict(ounet.state_dict())
    mb = mb.to(cuda_dev)
    phi_b, raw_b = xai.compute_time_shap_permutation(mb, sched, clfs["bf16"], x_T.to(cuda_dev), target, n_perm=M, seed=5, noise_seed=3)
    assert np.isfinite(phi_b).all() and abs(raw_b["efficiency_gap"]) < 1e-9
    assert abs(raw_b["prefix_values"][0, 0] - val_ref[0, 0]) < 5e-2                       # v(empty) does not depend on the sampler
