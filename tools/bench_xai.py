"""Time-SHAP sec/image (BASELINE config 4) and CSI batch-256 (config 5) on one GPU, with the CPU oracle beside it.

    python tools/bench_xai.py [--frames 1000] [--csi-batch 256] [--out gpurun_out/xai.json]
"""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from synt_isic_b200 import MelanomaClassifierAdaptive, xai  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=1000)
    ap.add_argument("--csi-batch", type=int, default=256)
    ap.add_argument("--out", default="gpurun_out/xai.json")
    ap.add_argument("--cpu-frames", type=int, default=16)
    ap.add_argument("--perm", type=int, default=4, help="permutations of the permutation Time-SHAP (0 = skip)")
    a = ap.parse_args()
    dev = torch.device("cuda:0")
    clf = MelanomaClassifierAdaptive(num_classes=7, pretrained=False, precision="bf16").to(dev).eval()
    g = torch.Generator().manual_seed(0)
    traj = torch.tanh(torch.randn(a.frames, 3, 128, 128, generator=g)).to(dev)
    res = {}

    def timed(fn, reps=5):
        fn(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps / 1e3

    t = timed(lambda: xai.compute_time_shap(clf, traj, list(range(a.frames)), 0))
    res["time_shap"] = {"frames": a.frames, "sec_per_image": t, "frames_per_s": a.frames / t,
                        "tflops": a.frames * 3.627e9 / t / 1e12}
    logits_t = timed(lambda: clf(traj))
    res["resnet18_logits"] = {"batch": a.frames, "sec": logits_t, "img_per_s": a.frames / logits_t}
    imgs = traj[: a.csi_batch].contiguous()
    masks = (torch.rand(a.csi_batch, 128, 128, generator=g) > 0.9).float().to(dev)
    noise = torch.randn(a.csi_batch, 3, 128, 128, generator=g).to(dev)
    kinds = ["noise", "blur", "shuffle", "zero", "mean"]
    tc = [i % 7 for i in range(a.csi_batch)]
    t = timed(lambda: xai.csi_batch(clf, imgs, masks, [k for k in kinds if k != "shuffle"], tc, noise=noise), reps=3)
    res["csi_batch"] = {"batch": a.csi_batch, "interventions": [k for k in kinds if k != "shuffle"], "sec": t,
                        "evals_per_s": a.csi_batch * 5 / t}
    attr_t = timed(lambda: xai.compute_shap_approximation(clf, imgs[:1], 0, n_samples=512), reps=3)
    res["patch_shap_512"] = {"sec_per_image": attr_t}
    # permutation Time-SHAP over the denoising steps (README.md:171-221): T = 50 steps, M permutations, 51 coalitions/batch
    if a.perm > 0:
        from synt_isic_b200 import DDPMScheduler, SUPPORTED_CONFIG, UNet2DModel
        model = UNet2DModel(precision="bf16", **SUPPORTED_CONFIG).to(dev)
        sched = DDPMScheduler(num_train_timesteps=1000, beta_schedule="squaredcos_cap_v2", prediction_type="epsilon")
        sched.set_timesteps(50)
        x_T = torch.randn(1, 3, 128, 128, generator=g).to(dev)
        xai.compute_time_shap_permutation(model, sched, clf, x_T, 0, n_perm=1, seed=1)          # warm-up (graph capture)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        phi, raw = xai.compute_time_shap_permutation(model, sched, clf, x_T, 0, n_perm=a.perm, seed=1)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        res["time_shap_permutation"] = {"steps": 50, "permutations": a.perm, "coalitions": a.perm * 51, "sec_per_image": dt,
                                        "unet_steps_per_s": a.perm * 51 * 50 / dt, "efficiency_gap": raw["efficiency_gap"]}
    # CPU oracle beside it (reference issues 2 forwards per frame at B=1)
    from oracle import xai as oxai
    from oracle.classifier import build_classifier
    torch.set_num_threads(os.cpu_count())
    oc = build_classifier()
    frames_cpu = [traj[i:i + 1].cpu() for i in range(a.cpu_frames)]
    oxai.time_shap(oc, frames_cpu[:2], [0, 1], 0)
    t0 = time.perf_counter()
    oxai.time_shap(oc, frames_cpu, list(range(a.cpu_frames)), 0)
    dt = time.perf_counter() - t0
    res["cpu_time_shap"] = {"frames_timed": a.cpu_frames, "sec_per_frame": dt / a.cpu_frames,
                            "sec_per_image_extrapolated": dt / a.cpu_frames * a.frames, "cores": os.cpu_count(),
                            "kind": "port (real torchvision resnet18, reference call pattern: 2 forwards/frame at B=1)"}
    print(json.dumps(res, indent=1))
    os.makedirs(os.path.dirname(a.out) or ".", exist_ok=True)
    json.dump(res, open(a.out, "w"), indent=1)


if __name__ == "__main__":
    main()
