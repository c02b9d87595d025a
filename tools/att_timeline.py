"""Per-CTA timeline of the persistent attention kernel (debug build).  On the GPU box:

    SYNT_EXTRA_NVCC_FLAGS=-DSYNT_ATT_TIMELINE_BUILD python -m synt_isic_b200.build --force
    SYNT_ATT_TIMELINE=gpurun_out/att_tl.bin python tools/att_bench.py; python tools/att_timeline.py gpurun_out/att_tl.bin
    python -m synt_isic_b200.build --force          # back to the product build

Record per CTA and softmax warp: smid, entry clock, then per item {end clock, cycles waiting for S tiles, for free P tiles,
for the accumulators (o_full), for the query tile (q_free), item id}."""
import sys
import numpy as np

a = np.fromfile(sys.argv[1], dtype=np.int64).reshape(-1, 8, 128)
ncta = a.shape[0]
smid = a[:, 0, 0]
entry = a[:, :, 1]
items = a[:, :, 8:8 + 20 * 6].reshape(ncta, 8, 20, 6)
n_items = (items[:, 0, :, 0] > 0).sum(1)
print(f"CTAs {ncta}, SMs used {len(set(smid.tolist()))}, CTAs per SM: {np.bincount(np.bincount(smid))[1:]} (index = count)")
print(f"items per CTA: min {n_items.min()} max {n_items.max()}")
t0 = entry.min()
end_all = items[:, :, :, 0].max()
print(f"kernel span {end_all - t0} clk; entry spread {entry.max() - t0}")
w = 0                                                     # warp 0 of warpgroup 0
first_end = items[:, w, 0, 0] - entry[:, w]
print(f"first item: entry -> end  mean {first_end.mean():.0f} clk")
dur = np.diff(items[:, w, :, 0], axis=1)
valid = items[:, w, 1:, 0] > 0
print(f"later items: duration mean {dur[valid].mean():.0f} min {dur[valid].min()} max {dur[valid].max()} clk")
for k, name in enumerate(["wait S tile", "wait free P tile", "wait accumulators (o_full)", "wait query tile (q_free)"]):
    v = items[:, w, 1:, 1 + k][valid]
    print(f"  {name:32s} mean {v.mean():8.0f} clk per item ({100 * v.mean() / dur[valid].mean():.1f}%)")
for wg, ww in (("wg0", 0), ("wg1", 4)):
    v = items[:, ww, 1:, 1:5][items[:, ww, 1:, 0] > 0]
    print(f"{wg}: waits per item  S {v[:, 0].mean():.0f}  P {v[:, 1].mean():.0f}  O {v[:, 2].mean():.0f}  Q {v[:, 3].mean():.0f}")
# are the two CTAs of an SM in phase?  item-end clocks modulo the mean item duration
sm0 = np.where(smid == smid[0])[0]
print("CTAs on SM", smid[0], ":", sm0.tolist())
for c in sm0:
    print("  cta", c, "entry", entry[c, 0] - t0, "item ends", (items[c, 0, :6, 0] - t0).tolist())
