"""Small fixed workload for ncu: a few eager sampling steps of the bf16 path at the bench batch.

    python tools/ncu_target.py [--batch 64] [--steps 2]
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from synt_isic_b200 import DDPMScheduler, SUPPORTED_CONFIG, UNet2DModel  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--steps", type=int, default=2)
    a = ap.parse_args()
    dev = torch.device("cuda:0")
    m = UNet2DModel(precision="bf16", **SUPPORTED_CONFIG).to(dev)
    s = DDPMScheduler(beta_schedule="squaredcos_cap_v2")
    s.set_timesteps(1000)
    x = torch.randn(a.batch, 3, 128, 128, device=dev)
    st = torch.cuda.Stream()
    with torch.cuda.stream(st):
        m.sample(x, s, step_begin=0, step_end=a.steps, use_graph=False)
        st.synchronize()
    print("ok", float(x.abs().mean()))


if __name__ == "__main__":
    main()
