"""Hardware experiment: are shifted (non-1024-B-aligned) SWIZZLE_128B views of a halo tile legal UMMA
operands?  Prints the error of conv_tc_halo for each descriptor/layout variant and its speed vs conv_tc."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import test_gpu_conv as T  # noqa: E402
from synt_isic_b200 import _lib  # noqa: E402

dev = torch.device("cuda:0")


def run(mode, B, H, W, Cin, Cout, iters=0):
    g = torch.Generator().manual_seed(0)
    x = torch.randn(B, Cin, H, W, generator=g)
    w = torch.randn(Cout, Cin, 3, 3, generator=g) / (Cin * 9) ** 0.5
    b = torch.randn(Cout, generator=g)
    q = lambda t: t.to(torch.bfloat16).float()
    ref = torch.nn.functional.conv2d(q(x).to(dev), q(w).to(dev), b.to(dev), padding=1)
    xin = x.permute(0, 2, 3, 1).contiguous().to(torch.bfloat16).to(dev)
    wp = T._pack(w).to(torch.bfloat16).to(dev)
    bd = b.to(dev)
    out = torch.zeros(B, H, W, Cout, dtype=torch.bfloat16, device=dev)
    call = lambda: _lib.check(_lib.lib().synt_debug_conv(mode, 1, xin.data_ptr(), B, H, W, Cin, 3, 1, 1, None, 0, None, 0, 1,
                                                         wp.data_ptr(), bd.data_ptr(), None, None, 0, out.data_ptr(), Cout,
                                                         _lib.current_stream_ptr()))
    call()
    torch.cuda.synchronize()
    got = out.float().permute(0, 3, 1, 2)
    err = ((got - ref).norm() / ref.norm()).item()
    ms = None
    if iters:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            call()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / iters
    return err, ms


for name, mode in (("conv_tc (9 loads)", 1), ("halo p16 boff=0", 2), ("halo p16 boff=dx", 3), ("halo p10 boff=0", 4),
                   ("halo p10 boff=addr", 5)):
    try:
        e1, _ = run(mode, 2, 32, 32, 64, 64)
        e2, ms = run(mode, 16, 128, 128, 64, 64, iters=20)
        fl = 2 * 16 * 128 * 128 * 64 * 576
        print(f"{name:22s} err(small)={e1:.3e} err(128^2)={e2:.3e}  {ms * 1e3:8.1f} us  {fl / ms / 1e9:7.1f} TFLOP/s", flush=True)
    except Exception as ex:
        print(f"{name:22s} FAILED: {ex}", flush=True)
