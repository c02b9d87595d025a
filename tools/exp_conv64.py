"""Where does the time of the Cout=64 128x128 layers go?  Times conv_tc2 (B=64, 128x128, 64->64, K=576) with
the epilogue / input features switched on one by one."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from synt_isic_b200 import _lib  # noqa: E402

dev = torch.device("cuda:0")


def run(B, H, W, Cin, Cout, gn, res, stats, K=3, iters=10):
    g = torch.Generator().manual_seed(0)
    x = (torch.randn(B, H, W, Cin, generator=g)).to(torch.bfloat16).to(dev)
    w = (torch.randn(Cout, K * K * Cin, generator=g) / (K * K * Cin) ** 0.5).to(torch.bfloat16).to(dev)
    b = torch.randn(Cout, generator=g).to(dev)
    ss = torch.stack([0.5 + torch.rand(B, Cin, generator=g), torch.randn(B, Cin, generator=g)], dim=2).contiguous().to(dev)
    r = torch.randn(B, H, W, Cout, generator=g).to(torch.bfloat16).to(dev) if res else None
    out = torch.empty(B, H, W, Cout, dtype=torch.bfloat16, device=dev)
    st = torch.zeros(B * 256 * Cout * 2, dtype=torch.float32, device=dev) if stats else None
    slots = C.c_int()
    ptr = lambda t: t.data_ptr() if t is not None else None
    call = lambda: _lib.check(_lib.lib().synt_debug_conv_gn(ptr(x), Cin, None, 0, ptr(ss) if gn else None, 2 if gn else 0, B, H, W, K,
                                                            None, 0, ptr(w), ptr(b), ptr(r), ptr(out), Cout, ptr(st), C.byref(slots),
                                                            _lib.current_stream_ptr()))
    call(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        call()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    fl = 2.0 * B * H * W * Cout * K * K * Cin
    return ms * 1e3, fl / ms / 1e9


SHAPES = ((64, 128, 128, 64, 64), (64, 128, 128, 128, 64), (64, 64, 64, 128, 128), (64, 32, 32, 256, 256), (64, 16, 16, 256, 256))
ITERS = int(os.environ.get("EXP_ITERS", "10"))
if os.environ.get("EXP_K1"):                 # the 1x1 projections of the attention blocks (qkv, out-projection)
    for shape, cfgs in (((64, 32, 32, 256, 768), ((0, 0, 0), (1, 0, 0))), ((64, 32, 32, 256, 256), ((0, 0, 0), (0, 1, 1)))):
        for gn, res, stats in cfgs:
            us, tf = run(*shape, gn, res, stats, K=1, iters=ITERS)
            print(f"1x1 {shape}  gn={gn} res={res} stats={stats}: {us:8.1f} us  {tf:7.1f} TFLOP/s", flush=True)
    sys.exit(0)
for shape in SHAPES:
    for gn, res, stats in ((0, 0, 0), (0, 1, 0), (0, 0, 1), (1, 0, 0), (1, 1, 1)):
        us, tf = run(*shape, gn, res, stats, iters=ITERS)
        print(f"{shape}  gn={gn} res={res} stats={stats}: {us:8.1f} us  {tf:7.1f} TFLOP/s", flush=True)
