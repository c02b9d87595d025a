"""Integrated Gradients (xai/XAI.py:1039-1085, 50 riemann_right steps) sec/attribution and the classifier's score +
input-gradient throughput on one GPU, with the CPU oracle (torch autograd through torchvision resnet18) beside it.

    python tools/bench_ig.py [--batch 64] [--out gpurun_out/ig.json]
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
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--out", default="gpurun_out/ig.json")
    ap.add_argument("--one", action="store_true",
                    help="one score + input-gradient pass at --batch between cudaProfilerStart/Stop (for `ncu --profile-from-start off`)")
    a = ap.parse_args()
    dev = torch.device("cuda:0")
    clf = MelanomaClassifierAdaptive(num_classes=7, pretrained=False, precision="bf16").to(dev).eval()
    g = torch.Generator().manual_seed(0)
    imgs = torch.tanh(torch.randn(a.batch, 3, 128, 128, generator=g)).to(dev)
    base = (torch.randn(1, 3, 128, 128, generator=g) * 0.1).to(dev)

    if a.one:
        for _ in range(2):
            clf.score_and_input_gradient(imgs, 0)
        torch.cuda.synchronize()
        torch.cuda.profiler.start()
        clf.score_and_input_gradient(imgs, 0)
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
        print("one pass done")
        return

    def timed(fn, reps=10):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps / 1e3

    res = {}
    t = timed(lambda: xai.compute_integrated_gradients(clf, imgs[:1], 0, n_steps=a.steps, baseline=base))
    res["integrated_gradients"] = {"n_steps": a.steps, "sec_per_attribution": t, "gradient_passes_per_s": a.steps / t}
    t = timed(lambda: clf.score_and_input_gradient(imgs, 0))
    # forward 3.627 GFLOP + data gradients of every conv but the stem's input side (~3.4 GFLOP algorithmic) per image
    res["score_and_input_gradient"] = {"batch": a.batch, "sec": t, "img_per_s": a.batch / t}
    t = timed(lambda: clf(imgs))
    res["logits_only"] = {"batch": a.batch, "sec": t, "img_per_s": a.batch / t}
    from oracle import xai as oxai
    from oracle.classifier import build_classifier
    torch.set_num_threads(os.cpu_count())
    oc = build_classifier()
    x_cpu, b_cpu = imgs[:1].cpu(), base.cpu()
    oxai.integrated_gradients(oc, x_cpu, 0, b_cpu, n_steps=5)
    t0 = time.perf_counter()
    oxai.integrated_gradients(oc, x_cpu, 0, b_cpu, n_steps=a.steps)
    res["cpu_integrated_gradients"] = {"sec_per_attribution": time.perf_counter() - t0, "cores": os.cpu_count(),
                                       "kind": "port (captum riemann_right restated; autograd through the real torchvision resnet18)"}
    print(json.dumps(res, indent=1))
    os.makedirs(os.path.dirname(a.out) or ".", exist_ok=True)
    json.dump(res, open(a.out, "w"), indent=1)


if __name__ == "__main__":
    main()
