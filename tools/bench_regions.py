"""select_regions_advanced (xai/XAI.py:1340-1451) maps/s on one GPU next to the numpy/scipy oracle on the host.

    python tools/bench_regions.py [--maps 512] [--out gpurun_out/regions.json]
"""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import scipy.ndimage as ndi  # noqa: E402
import torch  # noqa: E402

from synt_isic_b200 import xai  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--maps", type=int, default=512)
    ap.add_argument("--cpu-maps", type=int, default=32)
    ap.add_argument("--out", default="gpurun_out/regions.json")
    a = ap.parse_args()
    rng = np.random.default_rng(0)
    base = rng.standard_normal((a.maps, 3, 128, 128)).astype(np.float32)
    base = ndi.gaussian_filter(base, (0, 0, 2.0, 2.0)).astype(np.float32)
    dev = torch.device("cuda:0")
    x = torch.from_numpy(base).to(dev)
    for _ in range(3):
        xai.select_regions_batch(x, 10, "top", True, 8)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        xai.select_regions_batch(x, 10, "top", True, 8)
    e1.record(); torch.cuda.synchronize()
    t = e0.elapsed_time(e1) / 10 / 1e3
    from oracle import xai as oxai
    t0 = time.perf_counter()
    for i in range(a.cpu_maps):
        oxai.select_regions(base[i], 10, "top", True, 8)
    tc = (time.perf_counter() - t0) / a.cpu_maps
    res = {"maps": a.maps, "sec_per_launch": t, "maps_per_s": a.maps / t, "us_per_map": t / a.maps * 1e6,
           "cpu_sec_per_map": tc, "cpu_kind": "reference calls (numpy.percentile + scipy.ndimage), 1 core",
           "input_bytes_per_map": 3 * 128 * 128 * 4, "achieved_gb_s_input": a.maps * 3 * 128 * 128 * 4 * 3 / t / 1e9}
    print(json.dumps(res, indent=1))
    os.makedirs(os.path.dirname(a.out) or ".", exist_ok=True)
    json.dump(res, open(a.out, "w"), indent=1)


if __name__ == "__main__":
    main()
