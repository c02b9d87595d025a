"""Generates tests/golden/*.npz|json from the ORACLE (there are no reference-side golden vectors:
the reference has no tests, and `diffusers` is not installed here -- SURVEY.md section 8c).

    python tests/golden/make_golden.py

Fixtures are small (sub-sampled tensors + integer tables) so they travel with the repo:
  unet_eps.npz      eps[:, :, ::16, ::16] of the oracle UNet (class-0 weights) for seeded x at t in (980, 500, 0)
  ddpm_tables.json  timesteps (n=50, 1000, 7), sha256 of alphas_cumprod / coefficient tables (fp32 bytes),
                    the SURVEY.md A.3 known answers, md5 class seed offsets
  classifier.npz    logits of the oracle ResNet18 (seed-7 weights) for 4 seeded images, preprocess samples
  xai.npz           time-shap / patch-shap / intervention / cfi values on a seeded 6-frame trajectory
"""
import hashlib
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import xai as oxai  # noqa: E402
from oracle.classifier import build_classifier  # noqa: E402
from oracle.ddpm import DDPMSchedulerOracle  # noqa: E402
from oracle.unet2d import build_unet  # noqa: E402


def sha(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def main():
    torch.set_num_threads(os.cpu_count())
    # ---- UNet
    m = build_unet(0)
    g = torch.Generator().manual_seed(2024)
    x = torch.randn(2, 3, 128, 128, generator=g)
    out = {}
    with torch.no_grad():
        for t in (980, 500, 0):
            out[f"eps_t{t}"] = m(x, t).sample[:, :, ::16, ::16].numpy()
    out["param_count"] = np.array(sum(p.numel() for p in m.parameters()))
    np.savez(os.path.join(HERE, "unet_eps.npz"), **out)
    # ---- scheduler
    tables = {"known": {}}
    s = DDPMSchedulerOracle()
    tables["acp_sha256"] = sha(s.alphas_cumprod.numpy())
    tables["betas_sha256"] = sha(s.betas.numpy())
    for n in (50, 1000, 7):
        s.set_timesteps(n)
        coef = np.array([[float(v) for v in s.coefficients(int(t))] for t in s.timesteps], dtype=np.float32)
        coef[-1, 4] = 0.0                                   # t == 0: no noise is added
        tables[f"timesteps_{n}"] = s.timesteps.tolist()
        tables[f"coef_sha256_{n}"] = sha(coef)
    s.set_timesteps(50)
    tables["known"] = {
        "beta_0": float(s.betas[0]), "beta_500": float(s.betas[500]), "beta_998": float(s.betas[998]),
        "beta_999": float(s.betas[999]), "acp_0": float(s.alphas_cumprod[0]), "acp_500": float(s.alphas_cumprod[500]),
        "acp_980": float(s.alphas_cumprod[980]), "acp_999": float(s.alphas_cumprod[999]),
        "var_980_960": float(s.coefficients(980)[4] ** 2), "var_20_0": float(s.coefficients(20)[4] ** 2),
    }
    lin = DDPMSchedulerOracle(beta_schedule="linear")
    tables["linear_acp_sha256"] = sha(lin.alphas_cumprod.numpy())
    tables["md5_offsets"] = {c: int(hashlib.md5(c.encode()).hexdigest()[:8], 16) & 0x7FFFFFFF
                            for c in ["MEL", "NV", "BCC", "AKIEC", "BKL", "DF", "VASC"]}
    with open(os.path.join(HERE, "ddpm_tables.json"), "w") as f:
        json.dump(tables, f, indent=1)
    # ---- classifier
    c = build_classifier()
    imgs = torch.tanh(torch.randn(4, 3, 128, 128, generator=g) * 1.5)
    with torch.no_grad():
        logits = c(imgs).numpy()
        pre = c.preprocess_for_classifier(imgs)[:, :, ::28, ::28].numpy()
    np.savez(os.path.join(HERE, "classifier.npz"), logits=logits, preprocess_sub=pre)
    # ---- xai
    traj = [torch.tanh(torch.randn(1, 3, 128, 128, generator=g)) for _ in range(6)]
    imp, raw = oxai.time_shap(c, traj, list(range(6)), 0)
    masks = torch.rand(16, 8, 8, generator=g) > 0.5
    attr = oxai.patch_shap(c, traj[-1], 0, masks)
    mask = torch.zeros(128, 128)
    mask[40:80, 30:90] = 1
    blur, _ = oxai.intervene(traj[-1], mask, "blur")
    cs = oxai.causal_shift(c, traj[-1], blur, 0)
    np.savez(os.path.join(HERE, "xai.npz"), time_shap=imp, conf=raw["confidence_scores"], prob=raw["probability_scores"],
             patch_masks=masks.numpy(), patch_attr_sub=attr[0, 0, ::16, ::16].numpy(),
             blur_sub=blur[0, :, ::8, ::8].numpy(), cfi=np.array([cs["cfi"], cs["delta"], cs["kl_divergence"]]))
    print("golden fixtures written to", HERE)


if __name__ == "__main__":
    main()
