"""Multi-GPU parity on the box (skipped with fewer than 2 GPUs): the sharded classifier-evaluation paths and the sharded
data-set leg over NCCL give, bit for bit, what one rank computes alone (units are independent, weights replicated; the only
exchange is one gather / all_gather -- SURVEY.md section 8e)."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

_WORKER = r"""
import os, sys, numpy as np, torch, torch.distributed as dist
sys.path.insert(0, sys.argv[1])
from synt_isic_b200 import DDPMScheduler, MelanomaClassifierAdaptive, SUPPORTED_CONFIG, UNet2DModel, xai
from synt_isic_b200.dist import init_from_env, gather_images, partition
from synt_isic_b200.generator import to_uint8_tensor
rank, world, local = init_from_env("nccl")
dev = torch.device(f"cuda:{local}")
torch.cuda.set_device(dev)
torch.manual_seed(3)                                   # same random-init weights on every rank
clf = MelanomaClassifierAdaptive(num_classes=7, pretrained=False, precision="bf16").to(dev).eval()
g = torch.Generator().manual_seed(9)
frames = torch.tanh(torch.randn(301, 3, 128, 128, generator=g))
# Time-SHAP: device-resident and host-resident (pinned) trajectories, sharded vs alone
for src in (frames.to(dev), frames.pin_memory()):
    imp_s, raw_s = xai.compute_time_shap(clf, src, list(range(301)), 2, group=dist.group.WORLD)
    imp_1, raw_1 = xai.compute_time_shap(clf, src, list(range(301)), 2, group=None)
    assert np.array_equal(raw_s["confidence_scores"], raw_1["confidence_scores"]), "time-shap scores differ"
    assert np.array_equal(imp_s, imp_1)
# CSI batch: images split over the ranks
imgs = frames[:64].to(dev)
masks = (torch.rand(64, 128, 128, generator=g) > 0.9).float().to(dev)
noise = torch.randn(64, 3, 128, 128, generator=g).to(dev)
tc = [i % 7 for i in range(64)]
kinds = ["noise", "blur", "zero", "mean"]
a = xai.csi_batch(clf, imgs, masks, kinds, tc, noise=noise, group=dist.group.WORLD)
b = xai.csi_batch(clf, imgs, masks, kinds, tc, noise=noise, group=None)
for k in kinds:
    assert a[k].shape == (64,) and torch.equal(a[k], b[k]), k
# patch-SHAP coalitions sharded (one all_gather of the logits)
pm = torch.rand(33, 8, 8, generator=g) > 0.5
s1 = xai.compute_shap_approximation(clf, frames[:1].to(dev), 1, patch_masks=pm, group=dist.group.WORLD)
s2 = xai.compute_shap_approximation(clf, frames[:1].to(dev), 1, patch_masks=pm, group=None)
assert torch.equal(s1, s2)
# sampling units round-robin + ONE gather of the uint8 images to rank 0 == rank 0 sampling every unit itself
model = UNet2DModel(precision="bf16", **SUPPORTED_CONFIG).to(dev)
sched = DDPMScheduler(beta_schedule="squaredcos_cap_v2")
sched.set_timesteps(3)
units = list(range(2 * world))
def sample_unit(u):
    x = torch.randn(4, 3, 128, 128, generator=torch.Generator().manual_seed(100 + u)).to(dev)
    keys = torch.arange(4, dtype=torch.int64, device=dev) + 1000 * u
    model.sample(x, sched, seed=1, image_keys=keys)
    return to_uint8_tensor(x)
mine = partition(units, rank, world)
local_imgs = torch.cat([sample_unit(u) for u in mine])
allimg = gather_images(local_imgs, [len(partition(units, r, world)) * 4 for r in range(world)], dist.group.WORLD, dst=0)
if rank == 0:
    order = [u for r in range(world) for u in partition(units, r, world)]
    want = torch.cat([sample_unit(u) for u in order])
    assert allimg.shape == want.shape and torch.equal(allimg, want), "gathered images differ"
else:
    assert allimg is None
dist.barrier()
print(f"rank {rank} ok")
"""


@pytest.mark.parametrize("world", [2])
def test_sharded_paths_equal_one_rank(tmp_path, world):
    if not torch.cuda.is_available() or torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    script = tmp_path / "worker.py"
    script.write_text(_WORKER)
    port = 29500 + os.getpid() % 1000
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(port), str(script), ROOT]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert r.stdout.count(" ok") == world
