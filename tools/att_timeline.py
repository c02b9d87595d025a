"""Per-warp timeline of the attention kernel's softmax warps (debug build hook SYNT_ATT_TIMELINE): for CTA (1,1,1) of a
1024-token launch, clock64 at {unit start, S tile ready, P tile free, unit done} for every unit of every softmax warp.

    SYNT_EXTRA_NVCC_FLAGS=-DSYNT_ATT_TIMELINE_BUILD python -m synt_isic_b200.build --force   # instrumented build
    SYNT_ATT_TIMELINE=/tmp/att_tl.bin python tools/att_timeline.py
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from synt_isic_b200 import _lib  # noqa: E402

path = os.environ.setdefault("SYNT_ATT_TIMELINE", "/tmp/att_tl.bin")
if os.path.exists(path):
    os.unlink(path)
dev = torch.device("cuda:0")
B, N, C = 64, 1024, 256
g = torch.Generator().manual_seed(0)
qkv = (torch.randn(B, N, 3 * C, generator=g) * 0.7).to(torch.bfloat16).to(dev)
out = torch.empty(B, N, C, dtype=torch.bfloat16, device=dev)
for _ in range(3):
    _lib.check(_lib.lib().synt_debug_attention(1, 1, qkv.data_ptr(), B, N, C, out.data_ptr(), _lib.current_stream_ptr()))
torch.cuda.synchronize()
raw = np.fromfile(path, dtype=np.int64)
rec = raw.reshape(-1, 8, 64, 4)[-1]                  # last launch: [warp][slot][4]; slots 0,2,4,.. used
rec = rec[:, 0::2, :]                                # [8 warps][32 units][4]
t0 = rec[:, :, 0].min()
print("units per warp:", rec.shape[1], " CTA lifetime (clk):", int(rec[:, :, 3].max() - t0))
for w in range(8):
    r = rec[w] - t0
    wait_s = (r[:, 1] - r[:, 0])
    wait_p = (r[:, 2] - r[:, 1])
    work = (r[:, 3] - r[:, 2])
    gap = np.concatenate([[0], r[1:, 0] - r[:-1, 3]])
    print(f"warp {w + 4}: start {int(r[0, 0]):6d}  s_full wait avg {wait_s[1:].mean():7.1f} (first {int(wait_s[0])})  "
          f"max+p_free avg {wait_p[2:].mean():7.1f} (first chunk {wait_p[:2].mean():7.1f})  exp+store avg {work.mean():7.1f}  "
          f"unit period avg {np.diff(r[:, 0]).mean():7.1f}")
w0 = rec[0] - t0
print("warp 4 first units [start, S ready, P free, done]:")
print(w0[:8])
