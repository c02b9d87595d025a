"""Small fixed workload for ncu: ResNet18 logits of 512 images (argv[1] overrides) (bf16 path)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from synt_isic_b200 import MelanomaClassifierAdaptive  # noqa: E402

dev = torch.device("cuda:0")
clf = MelanomaClassifierAdaptive(num_classes=7, pretrained=False, precision="bf16").to(dev).eval()
x = torch.tanh(torch.randn(int(sys.argv[1]) if len(sys.argv) > 1 else 512, 3, 128, 128, device=dev))
for _ in range(2):
    y = clf(x)
torch.cuda.synchronize()
print("ok", float(y.abs().mean()))
