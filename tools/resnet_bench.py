"""ResNet18 logits throughput (bf16 path): 1000 frames (Time-SHAP sec/image) and the front-end / body / head split of 512."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from synt_isic_b200 import MelanomaClassifierAdaptive
dev = torch.device("cuda:0")
clf = MelanomaClassifierAdaptive(num_classes=7, pretrained=False, precision="bf16").to(dev).eval()
x = torch.tanh(torch.randn(1000, 3, 128, 128, device=dev))
for _ in range(3):
    y = clf(x)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    y = clf(x)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
print(f"1000 frames: {ms:.3f} ms  ({1000 / ms * 1e3:.0f} img/s, {1000 * 3.627e9 / (ms * 1e-3) / 1e12:.0f} TFLOP/s)  split of 512: {clf.profile_forward(x[:512])}  env "
      + " ".join(f"{k}={v}" for k, v in os.environ.items() if k.startswith("SYNT_")))
