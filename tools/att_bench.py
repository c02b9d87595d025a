"""Times the tcgen05 attention kernel alone at the step's shapes (B=64: five launches of N=1024 + one of N=256)."""
import sys, os, math, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from synt_isic_b200 import _lib
dev = torch.device("cuda:0")
B, C = 64, 256
L = _lib.lib()
res = {}
for N in (1024, 256):
    g = torch.Generator().manual_seed(N)
    qkv = (torch.randn(B, N, 3 * C, generator=g) * 0.7).to(torch.bfloat16).to(dev)
    out = torch.empty(B, N, C, dtype=torch.bfloat16, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    for _ in range(3):
        _lib.check(L.synt_debug_attention(1, 1, qkv.data_ptr(), B, N, C, out.data_ptr(), st))
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        _lib.check(L.synt_debug_attention(1, 1, qkv.data_ptr(), B, N, C, out.data_ptr(), st))
    e1.record(); torch.cuda.synchronize()
    res[N] = e0.elapsed_time(e1) / 10
step = 5 * res[1024] + res[256]
floor = B * 32 * (5 * 1024 ** 2 + 256 ** 2) / (16 * 148 * 1.965e9) * 1e3
print(f"attention us: N=1024 {res[1024]*1e3:.1f}  N=256 {res[256]*1e3:.1f}  per step {step:.3f} ms  (MUFU floor {floor:.3f} ms, frac {floor/step:.3f})  env "
      + " ".join(f"{k}={v}" for k, v in os.environ.items() if k.startswith("SYNT_ATT")))
