"""Layer-by-layer parity report of the CUDA path against the oracle (run under gpurun).

    python tools/gpu_check.py [--modes fp32,bf16] [--B 2] [--t 980] [--out gpurun_out/check.json]

For every module of the UNet (diffusers module paths) and the ResNet18 it prints the rel-L2 error
of the CUDA output against the oracle's forward-hook capture on the same input and weights.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch  # noqa: E402


def rel(a, b):
    return ((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30)).item()


def unet_taps():
    taps = ["conv_in"]
    for i in range(4):
        for j in range(2):
            taps.append(f"down_blocks.{i}.resnets.{j}")
            if i == 2:
                taps.append(f"down_blocks.{i}.attentions.{j}")
        if i != 3:
            taps.append(f"down_blocks.{i}.downsamplers.0")
    taps += ["mid_block.resnets.0", "mid_block.attentions.0", "mid_block.resnets.1"]
    for i in range(4):
        for j in range(3):
            taps.append(f"up_blocks.{i}.resnets.{j}")
            if i == 1:
                taps.append(f"up_blocks.{i}.attentions.{j}")
        if i != 3:
            taps.append(f"up_blocks.{i}.upsamplers.0")
    taps.append("conv_out")
    return taps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--modes", default="fp32,bf16")
    ap.add_argument("--B", type=int, default=2)
    ap.add_argument("--t", type=int, default=980)
    ap.add_argument("--out", default="gpurun_out/check.json")
    ap.add_argument("--skip-resnet", action="store_true")
    ap.add_argument("--skip-unet", action="store_true")
    args = ap.parse_args()
    from oracle.classifier import build_classifier
    from oracle.unet2d import build_unet
    from synt_isic_b200 import MelanomaClassifierAdaptive, SUPPORTED_CONFIG, UNet2DModel

    dev = torch.device("cuda:0")
    report = {"unet": {}, "resnet": {}}
    g = torch.Generator().manual_seed(123)
    if not args.skip_unet:
        oracle = build_unet(0)
        x = torch.randn(args.B, 3, 128, 128, generator=g)
        captured = {}
        mods = dict(oracle.named_modules())
        hooks = []
        for tap in unet_taps():
            hooks.append(mods[tap].register_forward_hook(
                lambda m, i, o, tap=tap: captured.__setitem__(tap, (o.sample if hasattr(o, "sample") else o).detach())))
        t0 = time.time()
        with torch.no_grad():
            ref_eps = oracle(x, args.t).sample
        print(f"oracle forward {time.time() - t0:.2f}s", flush=True)
        for h in hooks:
            h.remove()
        for mode in args.modes.split(","):
            model = UNet2DModel(precision=mode, **SUPPORTED_CONFIG)
            model.load_state_dict(oracle.state_dict())
            model = model.to(dev)
            xs = x.to(dev)
            rep = {}
            try:
                eps = model(xs, args.t).sample
                torch.cuda.synchronize()
                rep["eps"] = rel(eps.cpu(), ref_eps)
                print(f"[{mode}] eps rel-L2 = {rep['eps']:.3e}", flush=True)
            except Exception as e:
                print(f"[{mode}] forward FAILED: {e}", flush=True)
                rep["error"] = str(e)
                report["unet"][mode] = rep
                break
            for tap in unet_taps():
                try:
                    got = model.debug_tap(xs, args.t, tap).cpu()
                    torch.cuda.synchronize()
                    r = rel(got, captured[tap])
                except Exception as e:
                    r = f"ERR {e}"
                rep[tap] = r
                print(f"[{mode}] {tap:34s} {r if isinstance(r, str) else format(r, '.3e')}", flush=True)
            report["unet"][mode] = rep
            del model
    if not args.skip_resnet:
        oc = build_classifier()
        imgs = torch.tanh(torch.randn(4, 3, 128, 128, generator=g))
        caps = {}
        with torch.no_grad():
            pre = oc.preprocess_for_classifier(imgs)
            caps["preprocess"] = pre
            m = oc.model
            y = m.relu(m.bn1(m.conv1(pre))); caps["relu"] = y
            y = m.maxpool(y); caps["maxpool"] = y
            for li, layer in enumerate([m.layer1, m.layer2, m.layer3, m.layer4]):
                for bj, blk in enumerate(layer):
                    y = blk(y); caps[f"layer{li + 1}.{bj}"] = y
            ref_logits = oc(imgs)
        for mode in args.modes.split(","):
            clf = MelanomaClassifierAdaptive(num_classes=7, pretrained=False, precision=mode)
            clf.model.load_state_dict(oc.model.state_dict())
            clf = clf.to(dev).eval()
            rep = {}
            try:
                got = clf(imgs.to(dev)).cpu()
                rep["logits"] = rel(got, ref_logits)
                print(f"[{mode}] resnet logits rel-L2 = {rep['logits']:.3e}", flush=True)
            except Exception as e:
                print(f"[{mode}] resnet FAILED: {e}", flush=True)
                rep["error"] = str(e)
                report["resnet"][mode] = rep
                break
            for tap, ref in caps.items():
                try:
                    r = rel(clf.debug_tap(imgs.to(dev), tap).cpu(), ref)
                except Exception as e:
                    r = f"ERR {e}"
                rep[tap] = r
                print(f"[{mode}] resnet {tap:20s} {r if isinstance(r, str) else format(r, '.3e')}", flush=True)
            report["resnet"][mode] = rep
    os.makedirs(os.path.dirname(args.out) or ".", exist_ok=True)
    with open(args.out, "w") as f:
        json.dump(report, f, indent=1)


if __name__ == "__main__":
    main()
