"""Statistical validation (xai/XAI.py:1708-2005): the resampling loops on the GPU against the reference's numpy loops.

CPU part: the restated numpy summation order equals ``np.sum`` bit for bit, and the index arrays of
``draw_bootstrap_indices`` / ``draw_permutations`` replay the reference's global-RNG stream (applying them with plain numpy
gives exactly the replicate differences of the reference's loops).  GPU part: with those indices injected the CUDA
replicates are bit-identical to the reference's; with the in-kernel Philox stream the confidence interval and the p-value
agree within Monte-Carlo error, and the full result dictionary matches scipy's closed-form tests."""
import numpy as np
import pytest

from oracle import xai as oxai


def _samples(seed=0, n1=24, n2=19):
    rng = np.random.default_rng(seed)
    return rng.normal(0.8, 0.5, n1), rng.normal(0.3, 0.7, n2)


def test_numpy_summation_order_restated():
    rng = np.random.default_rng(1)
    for n in list(range(1, 200)) + [256, 257, 1000, 4099]:
        a = rng.standard_normal(n) * 10.0 ** rng.uniform(-3, 3)
        assert oxai.numpy_pairwise_sum(a) == float(np.sum(a)), n


def test_index_replay_equals_the_reference_loops():
    from synt_isic_b200 import xai                                   # host-side helpers only (no GPU call)
    top, bot = _samples()
    np.random.seed(123)
    boot_ref, perm_ref = oxai.bootstrap_and_permutation(top, bot, n_bootstrap=50, n_permutations=80)
    np.random.seed(123)
    it, ib = xai.draw_bootstrap_indices(len(top), len(bot), 50)
    perms = xai.draw_permutations(len(top) + len(bot), 80)
    boot = np.array([np.mean(top[it[b]]) - np.mean(bot[ib[b]]) for b in range(50)])
    comb = np.concatenate([top, bot])
    perm = np.array([np.mean(comb[perms[k]][:len(top)]) - np.mean(comb[perms[k]][len(top):]) for k in range(80)])
    assert np.array_equal(boot, boot_ref) and np.array_equal(perm, perm_ref)


@pytest.mark.gpu
@pytest.mark.parametrize("n1,n2", [(24, 19), (6, 6), (150, 131), (3, 2)])
def test_gpu_replicates_bit_identical_with_injected_draws(cuda_dev, n1, n2):
    from synt_isic_b200 import xai
    top, bot = _samples(n1 + n2, n1, n2)
    np.random.seed(7)
    boot_ref, perm_ref = oxai.bootstrap_and_permutation(top, bot, n_bootstrap=300, n_permutations=700)
    np.random.seed(7)
    idx = xai.draw_bootstrap_indices(n1, n2, 300)
    perms = xai.draw_permutations(n1 + n2, 700)
    boot, perm = xai.resampled_mean_differences(top, bot, 300, 700, device=str(cuda_dev), bootstrap_indices=idx, permutations=perms)
    assert boot.dtype == np.float64 and np.array_equal(boot, boot_ref)
    assert np.array_equal(perm, perm_ref)


@pytest.mark.gpu
def test_gpu_statistical_validation_dictionary(cuda_dev):
    from scipy import stats
    from synt_isic_b200 import xai
    top, bot = _samples(3, 40, 35)
    res = xai.statistical_validation_comprehensive(top, bot, device=str(cuda_dev), seed=11)
    assert set(res) == {"descriptive_statistics", "parametric_tests", "nonparametric_tests", "effect_sizes", "bootstrap_analysis",
                        "permutation_analysis", "normality_tests", "variance_tests", "significance_consensus",
                        "overall_conclusion", "metadata"}
    assert res["parametric_tests"]["welch_t_test"]["p_value"] == stats.ttest_ind(top, bot, equal_var=False)[1]
    assert res["nonparametric_tests"]["mann_whitney_u"]["statistic"] == stats.mannwhitneyu(top, bot, alternative="two-sided")[0]
    b, p = res["bootstrap_analysis"], res["permutation_analysis"]
    assert b["bootstrap_diffs"].shape == (1000,) and p["permuted_differences"].shape == (10000,)
    # in-kernel Philox draws against the reference's numpy loops: same estimands within Monte-Carlo error
    np.random.seed(5)
    boot_ref, perm_ref = oxai.bootstrap_and_permutation(top, bot)
    se = boot_ref.std() / np.sqrt(1000)
    assert abs(b["mean_diff"] - boot_ref.mean()) < 6 * se
    assert abs(b["ci_lower"] - np.percentile(boot_ref, 5)) < 0.05 and abs(b["ci_upper"] - np.percentile(boot_ref, 95)) < 0.05
    obs = np.mean(top) - np.mean(bot)
    p_ref = np.mean(np.abs(perm_ref) >= abs(obs))
    assert abs(p["p_value"] - p_ref) < 0.01 + 4 * np.sqrt(max(p_ref, 1e-4) / 10000)
    assert abs(p["permuted_differences"].mean()) < 0.01                      # a permutation null is centred on 0
    assert abs(p["permuted_differences"].std() - perm_ref.std()) < 0.01
    assert res["overall_conclusion"]["significant"] in (True, False) and res["overall_conclusion"]["total_tests_count"] == 4
    # two different seeds give different draws, the same seed the same
    a1 = xai.resampled_mean_differences(top, bot, 100, 100, device=str(cuda_dev), seed=1)
    a2 = xai.resampled_mean_differences(top, bot, 100, 100, device=str(cuda_dev), seed=1)
    a3 = xai.resampled_mean_differences(top, bot, 100, 100, device=str(cuda_dev), seed=2)
    assert np.array_equal(a1[0], a2[0]) and np.array_equal(a1[1], a2[1]) and not np.array_equal(a1[0], a3[0])
