"""Per-launch timing of one sampling step (CUDA-event pairs), grouped by GEMM shape.

    python tools/profile_layers.py [--batch 64] [--micro-batch 16] [--out gpurun_out/layers.txt]
"""
import argparse
import collections
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from synt_isic_b200 import DDPMScheduler, SUPPORTED_CONFIG, UNet2DModel  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--micro-batch", type=int, default=0)
    ap.add_argument("--out", default="gpurun_out/layers.txt")
    a = ap.parse_args()
    dev = torch.device("cuda:0")
    m = UNet2DModel(precision="bf16", **SUPPORTED_CONFIG).to(dev)
    s = DDPMScheduler(beta_schedule="squaredcos_cap_v2")
    s.set_timesteps(1000)
    x = torch.randn(a.batch, 3, 128, 128, device=dev)
    st = torch.cuda.Stream()
    with torch.cuda.stream(st):
        m.sample(x, s, step_begin=0, step_end=3, micro_batch=a.micro_batch)
        prof = m.profile_step(x, s, micro_batch=a.micro_batch)
        st.synchronize()
    recs = m.profile_records()
    agg = collections.OrderedDict()
    for cat, ms, fl, M, N, K in recs:
        key = (cat, M, N, K)
        e = agg.setdefault(key, [0, 0.0, 0.0])
        e[0] += 1; e[1] += ms; e[2] += fl
    lines = [f"{'category':16s} {'M':>8s} {'N':>5s} {'K':>5s} {'n':>4s} {'ms':>9s} {'us/launch':>10s} {'TFLOP/s':>8s}"]
    for (cat, M, N, K), (n, ms, fl) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        tf = fl / (ms * 1e-3) / 1e12 if ms > 0 and fl > 0 else 0.0
        lines.append(f"{cat:16s} {M:8d} {N:5d} {K:5d} {n:4d} {ms:9.3f} {1e3 * ms / n:10.1f} {tf:8.1f}")
    lines.append("")
    for k, v in prof.items():
        lines.append(f"{k:18s} {v['ms']:9.3f} ms  {v['launches']:4d} launches")
    lines.append(f"total {sum(v['ms'] for v in prof.values()):.3f} ms")
    txt = "\n".join(lines)
    print(txt)
    os.makedirs(os.path.dirname(a.out) or ".", exist_ok=True)
    open(a.out, "w").write(txt + "\n")


if __name__ == "__main__":
    main()
