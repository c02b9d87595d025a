"""TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

``MelanomaClassifierAdaptive`` restated from ``xai/XAI.py:357-471`` of the reference on
top of the REAL torchvision ``resnet18`` (torchvision/models/resnet.py) and the REAL
``F.interpolate(..., mode='bilinear', align_corners=False, antialias=True)`` call, so the
arithmetic at this boundary is the third-party code itself (pinned by construction).
Built the way ``xai/xai_integration.py:79`` builds it: 7 outputs; ImageNet weights are
not downloadable offline so weights are random-init.
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F
from torchvision import models

CLASS_NAMES = ["MEL", "NV", "BCC", "AKIEC", "BKL", "DF", "VASC"]   # XAI.py:196
NUM_CLASSES = 7
CLASSIFIER_IMAGE_SIZE = 224                                        # XAI.py:192
IMAGENET_MEAN = (0.485, 0.456, 0.406)                              # XAI.py:426
IMAGENET_STD = (0.229, 0.224, 0.225)                               # XAI.py:427
SCORE_EPS = 1e-8                                                   # XAI.py:459


class ClassifierOracle(nn.Module):
    def __init__(self, num_classes: int = NUM_CLASSES):
        super().__init__()
        self.num_classes = num_classes
        self.model = models.resnet18(weights=None)                 # XAI.py:389 (offline)
        self.model.fc = nn.Linear(self.model.fc.in_features, num_classes)   # XAI.py:393-394
        self.architecture = "resnet18"

    def preprocess_for_classifier(self, x):                        # XAI.py:399-431
        x = torch.clamp((x + 1.0) / 2.0, 0, 1)
        if x.shape[-1] != CLASSIFIER_IMAGE_SIZE or x.shape[-2] != CLASSIFIER_IMAGE_SIZE:
            x = F.interpolate(x, size=(CLASSIFIER_IMAGE_SIZE, CLASSIFIER_IMAGE_SIZE),
                              mode="bilinear", align_corners=False, antialias=True)
        mean = torch.tensor(IMAGENET_MEAN, dtype=x.dtype, device=x.device).view(1, 3, 1, 1)
        std = torch.tensor(IMAGENET_STD, dtype=x.dtype, device=x.device).view(1, 3, 1, 1)
        return (x - mean) / std                                    # transforms.Normalize

    def forward(self, x):                                          # XAI.py:433-436
        return self.model(self.preprocess_for_classifier(x))

    def get_probabilities(self, x):                                # XAI.py:438-441
        return F.softmax(self.forward(x), dim=1)

    def get_per_class_score(self, x, target_class):                # XAI.py:443-459
        return torch.log(self.get_probabilities(x)[:, target_class] + SCORE_EPS)

    def predict(self, x):                                          # XAI.py:461-465
        with torch.no_grad():
            return torch.argmax(self.forward(x), dim=1)

    def get_confidence(self, x, target_class):                     # XAI.py:467-471
        with torch.no_grad():
            return self.get_probabilities(x)[:, target_class]


def build_classifier(seed: int = 7) -> ClassifierOracle:
    """Random-init ResNet18 (``manual_seed(7)``) with BN running stats randomised
    (mean 0.1 N, var U(0.5,1.5)) and BN affine perturbed, so BN folding is exercised."""
    state = torch.random.get_rng_state()
    torch.manual_seed(seed)
    try:
        m = ClassifierOracle()
    finally:
        torch.random.set_rng_state(state)
    g = torch.Generator().manual_seed(seed + 1)
    with torch.no_grad():
        for mod in m.modules():
            if isinstance(mod, nn.BatchNorm2d):
                mod.running_mean.copy_(0.1 * torch.randn(mod.running_mean.shape, generator=g))
                mod.running_var.copy_(0.5 + torch.rand(mod.running_var.shape, generator=g))
                mod.weight.copy_(1.0 + 0.1 * torch.randn(mod.weight.shape, generator=g))
                mod.bias.copy_(0.1 * torch.randn(mod.bias.shape, generator=g))
    return m.eval()
