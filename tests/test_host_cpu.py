"""CPU: host-side logic of the product package and the C-ABI surface (no compute calls)."""
import ctypes
import hashlib
import json
import os
import re
import subprocess
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = os.path.join(ROOT, "tests", "golden")


def test_library_exports_every_header_symbol():
    from synt_isic_b200 import _lib
    hdr = open(os.path.join(ROOT, "include", "synt_isic.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(synt_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 25
    L = ctypes.CDLL(str(_lib.LIB_PATH))
    for name in declared:
        assert hasattr(L, name), f"{name} declared in include/synt_isic.h but not exported"
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    assert b"sm_100a" in _lib.lib().synt_version()


def test_unet_manifest_equals_oracle_state_dict():
    from oracle.unet2d import build_unet
    from synt_isic_b200 import _lib
    man = _lib.unet_manifest()
    sd = build_unet(0).state_dict()
    assert {n for n, _, _ in man} == set(sd)
    assert all(sd[n].numel() == k for n, k, _ in man)
    assert _lib.lib().synt_unet_total_param_count() == 25_304_963
    offs = [o for _, _, o in man]
    assert offs == sorted(offs) and offs[0] == 0


def test_resnet_manifest_equals_torchvision_state_dict():
    from oracle.classifier import build_classifier
    from synt_isic_b200 import _lib
    sd = {k: v for k, v in build_classifier().model.state_dict().items() if not k.endswith("num_batches_tracked")}
    man = _lib.resnet18_manifest(7)
    assert {n for n, _, _ in man} == set(sd)
    assert all(sd[n].numel() == k for n, k, _ in man)
    assert sum(k for _, k, _ in man) == 11_180_103 + sum(v.numel() for k, v in sd.items() if "running" in k)


def test_dropin_unet_state_dict_roundtrip_and_cpu_refusal():
    from oracle.unet2d import build_unet
    from synt_isic_b200 import SUPPORTED_CONFIG, UNet2DModel
    o = build_unet(1)
    m = UNet2DModel(**SUPPORTED_CONFIG)
    res = m.load_state_dict(o.state_dict(), strict=True)
    assert not res.missing_keys and not res.unexpected_keys
    assert all(torch.equal(m.state_dict()[k], v) for k, v in o.state_dict().items())
    assert not m.training and sum(p.numel() for p in m.parameters()) == 25_304_963
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m(torch.zeros(1, 3, 128, 128), 10)
    with pytest.raises(NotImplementedError):
        UNet2DModel(block_out_channels=(32, 64, 128, 128))
    # pre-0.15 attention key names are accepted
    old = {k.replace(".to_q.", ".query.").replace(".to_k.", ".key.").replace(".to_v.", ".value.")
            .replace(".to_out.0.", ".proj_attn."): v for k, v in o.state_dict().items()}
    assert not m.load_state_dict(old, strict=True).missing_keys
    packed = m._packed_params()
    assert packed.shape == (25_304_963,) and np.isfinite(packed).all()


def test_timestep_argument_forms():
    from synt_isic_b200 import UNet2DModel
    f = UNet2DModel._timestep_int
    assert f(980) == 980 and f(torch.tensor(980)) == 980 and f(torch.tensor([5])) == 5
    assert f(torch.tensor([7, 7, 7])) == 7
    with pytest.raises(NotImplementedError):
        f(torch.tensor([1, 2]))


def test_scheduler_dropin_bit_exact_vs_oracle():
    from oracle.ddpm import DDPMSchedulerOracle
    from synt_isic_b200 import DDPMScheduler
    tab = json.load(open(os.path.join(G, "ddpm_tables.json")))
    for sched in ("squaredcos_cap_v2", "linear"):
        for n in (None, 50, 1000, 7, 1):
            s, o = DDPMScheduler(beta_schedule=sched), DDPMSchedulerOracle(beta_schedule=sched)
            if n:
                s.set_timesteps(n)
                o.set_timesteps(n)
            assert s.timesteps.dtype == torch.int64
            assert torch.equal(s.timesteps, o.timesteps)                      # timestep indexing: bit-exact
            assert torch.equal(s.alphas_cumprod, o.alphas_cumprod)
            for t in o.timesteps.tolist():
                co = np.array([float(v) for v in o.coefficients(t)], dtype=np.float32)
                if t == 0:
                    co[4] = 0
                assert np.array_equal(co, s.coefficients(t)), (sched, n, t)
    s = DDPMScheduler(num_train_timesteps=1000, beta_schedule="squaredcos_cap_v2", prediction_type="epsilon")
    s.set_timesteps(50)
    assert s.timesteps.tolist() == tab["timesteps_50"]
    assert [int(t) for t in s.timesteps][:2] == [980, 960]                    # iterable of 0-d int64 tensors
    assert hashlib.sha256(s._coef.tobytes()).hexdigest() == tab["coef_sha256_50"]
    with pytest.raises(ValueError):
        s.set_timesteps(1001)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        s.step(torch.zeros(1, 3, 8, 8), 980, torch.zeros(1, 3, 8, 8))


def test_c_ddpm_tables_twin():
    from synt_isic_b200 import DDPMScheduler
    tab = json.load(open(os.path.join(G, "ddpm_tables.json")))
    for n in (50, 1000, 7):
        ts, coef, acp = DDPMScheduler.c_tables(1000, "squaredcos_cap_v2", 1e-4, 0.02, n)
        assert ts.tolist() == tab[f"timesteps_{n}"]                           # bit-exact indices
        assert hashlib.sha256(acp.tobytes()).hexdigest() == tab["acp_sha256"]  # bit-exact alphas_cumprod
        s = DDPMScheduler(beta_schedule="squaredcos_cap_v2")
        s.set_timesteps(n)
        np.testing.assert_allclose(coef, s._coef, rtol=3e-7, atol=0)         # torch's sqrt is 1 ulp off at places
    ts, coef, acp = DDPMScheduler.c_tables(1000, "linear", 1e-4, 0.02, 1000)
    s = DDPMScheduler(beta_schedule="linear")
    np.testing.assert_allclose(acp, s.alphas_cumprod.numpy(), rtol=1e-4)


def test_seed_algebra_and_filenames():
    from synt_isic_b200 import generator as gen
    tab = json.load(open(os.path.join(G, "ddpm_tables.json")))
    for c, off in tab["md5_offsets"].items():
        assert gen.class_seed_offset(c) == off
    assert gen.image_seed(42, "MEL", 0) == (42 + 2133561680) & 0x7FFFFFFF
    assert gen.image_seed(2 ** 31 - 1, "MEL", 5) == (2 ** 31 - 1 + 2133561680 + 5) & 0x7FFFFFFF
    assert gen.isic_filename(12) == "ISIC_0000012.png"
    x = torch.randn(1, 3, 128, 128, generator=torch.Generator().manual_seed(1))
    assert gen.noise_hash(x) == hashlib.sha256(x.numpy().tobytes()).hexdigest()[:16]
    img = (np.random.RandomState(0).rand(16, 16, 3) * 255).astype(np.uint8)
    assert np.array_equal(gen.color_postprocess(img, None), img)
    out = gen.color_postprocess(img, {"rgb": {"mean": [100, 110, 120], "std": [40, 40, 40]}})
    assert out.dtype == np.uint8 and out.shape == img.shape


def test_classifier_dropin_container():
    from synt_isic_b200 import MelanomaClassifierAdaptive
    c = MelanomaClassifierAdaptive(num_classes=7, architecture="auto", pretrained=True)
    assert c.model.fc.out_features == 7 and not c.training
    assert c.model.layer4[-1].conv2.weight.shape == (512, 512, 3, 3)       # Grad-CAM handle, XAI.py:2946
    assert len(list(c.named_parameters())) == 62
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        c(torch.zeros(1, 3, 128, 128))
    k0 = c._version_key()
    with torch.no_grad():
        next(c.parameters()).add_(1.0)                                     # XAI.py:2055-2059 mutates in place
    assert c._version_key() != k0


def test_shard_bounds_and_partition():
    from synt_isic_b200.dist import partition, shard_bounds
    for n in (0, 1, 7, 64, 513):
        for w in (1, 2, 3, 8):
            b = [shard_bounds(n, r, w) for r in range(w)]
            assert b[0][0] == 0 and b[-1][1] == n and all(b[i][1] == b[i + 1][0] for i in range(w - 1))
            assert max(h - l for l, h in b) - min(h - l for l, h in b) <= 1
    units = list(range(125))                                               # config 3: 125 batches of 64
    got = sorted(sum((partition(units, r, 8) for r in range(8)), []))
    assert got == units and max(len(partition(units, r, 8)) for r in range(8)) == 16


_WORKER = r"""
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, sys.argv[1])
from synt_isic_b200.dist import init_from_env, sharded_eval, gather_images, max_over_ranks, shard_bounds
rank, world, _ = init_from_env("gloo")
calls = []
def fn(x):                      # stand-in for the classifier: row-wise, deterministic
    calls.append(x.shape[0])
    return torch.stack([x.sum(dim=(1, 2, 3)), x.mean(dim=(1, 2, 3)) * 3], dim=1)
g = torch.Generator().manual_seed(5)
items = torch.randn(13, 3, 4, 4, generator=g)
out = sharded_eval(fn, items, dist.group.WORLD, chunk=4)
ref = torch.stack([items.sum(dim=(1, 2, 3)), items.mean(dim=(1, 2, 3)) * 3], dim=1)
assert torch.allclose(out, ref, atol=1e-6), (out - ref).abs().max()
lo, hi = shard_bounds(13, rank, world)
assert sum(calls) == hi - lo                     # each rank evaluated only its slice
counts = [3, 2]
mine = torch.full((counts[rank], 2, 2, 3), rank + 1, dtype=torch.uint8)
allimg = gather_images(mine, counts, dist.group.WORLD, dst=0)
if rank == 0:
    assert allimg.shape == (5, 2, 2, 3) and allimg[:3].eq(1).all() and allimg[3:].eq(2).all()
else:
    assert allimg is None
assert max_over_ranks(float(rank + 1), "cpu", dist.group.WORLD) == 2.0
solo = sharded_eval(fn, items, None, chunk=4)          # group=None: this rank alone, no collective
assert torch.allclose(solo, ref, atol=1e-6)
# bulk-dataset driver (SURVEY 8f row 1): units round-robin over the ranks, ONE gather_object of the rows, CSV on rank 0
import numpy as np, csv, tempfile
from synt_isic_b200 import bulk
class FakeGen:
    base_seed = 42; color_statistics = {}; stop_requested = False
    def generate_batch(self, class_name, seeds):
        return np.zeros((len(seeds), 128, 128, 3), np.uint8), None, ["h"] * len(seeds)
out_dir = sys.argv[2]
written = []
res = bulk.generate_dataset(FakeGen(), [("MEL", 70), ("DF", 5), ("NV", 130)], out_dir, layout="flat", postprocess=False,
                            batch_size=64, group=dist.group.WORLD, save_fn=lambda img, fp, c, s, h: written.append(fp.name))
units = bulk.plan_units([("MEL", 70), ("DF", 5), ("NV", 130)], 64, "flat")
assert len(written) == sum(u.count for i, u in enumerate(units) if i % world == rank)
if rank == 0:
    assert res["total"] == 205 and res["generated"] == {"MEL": 70, "DF": 5, "NV": 130}
    rows = list(csv.reader(open(res["files"]["ground_truth_csv"])))
    assert len(rows) == 206 and rows[1][0] == "ISIC_0034321.jpg" and rows[-1][0] == "ISIC_0034525.jpg"
    assert [r[0] for r in rows[1:]] == sorted(r[0] for r in rows[1:])          # same file as a single-rank run would write
# Integrated Gradients over a stack of frames, images split over the ranks (closed-form stand-ins for the CUDA entry points)
import contextlib, ctypes as C
from synt_isic_b200 import _lib, xai
def arr(ptr, n):
    return np.ctypeslib.as_array((C.c_float * n).from_address(ptr))
class FakeLib:
    def synt_ig_interpolate(self, x, b, n, per, out, st):
        X, B, O = arr(x, per), arr(b, per), arr(out, n * per).reshape(n, per)
        for k in range(n):
            O[k] = B + np.float32((k + 1) / n) * (X - B)
        return 0
    def synt_ig_reduce(self, g, x, b, n, per, out, st):
        arr(out, per)[:] = (arr(x, per) - arr(b, per)) * arr(g, n * per).reshape(n, per).sum(0) / n
        return 0
class FakeClassifier:
    seen = 0
    def parameters(self):
        yield torch.zeros(1)
    def score_and_input_gradient(self, pts, target):
        FakeClassifier.seen += pts.shape[0]
        return pts.flatten(1).pow(2).sum(1), 2 * pts
_lib.lib = lambda: FakeLib()
_lib.current_stream_ptr = lambda: 0
torch.cuda.device = lambda d: contextlib.nullcontext()
gi = torch.Generator().manual_seed(11)
frames = torch.randn(5, 3, 128, 128, generator=gi)
base = torch.randn(1, 3, 128, 128, generator=gi) * 0.1
ig = xai.compute_integrated_gradients_batch(FakeClassifier(), frames, 0, n_steps=4, baselines=base, images_per_pass=2,
                                            group=dist.group.WORLD)
want = (frames - base) * sum(2 * (base + a * (frames - base)) for a in (0.25, 0.5, 0.75, 1.0)) / 4
assert ig.shape == (5, 3, 128, 128) and torch.allclose(ig, want, atol=1e-5)
lo, hi = shard_bounds(5, rank, world)
assert FakeClassifier.seen == (hi - lo) * 4                      # this rank differentiated only its own frames
dist.barrier()
sys.stdout.write(f"RANK_OK {rank}\n")          # ONE write per rank: both ranks share the pipe, print() may interleave its pieces
sys.stdout.flush()
"""


def test_world_size_2_gloo(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(_WORKER)
    port = 29500 + (os.getpid() % 400)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr",
           "127.0.0.1", "--master-port", str(port), str(script), ROOT, str(tmp_path / "dataset")]
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="", OMP_NUM_THREADS="1")
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=240, env=env)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "RANK_OK 0" in r.stdout and "RANK_OK 1" in r.stdout


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "synt_isic_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            src = open(os.path.join(pkg, fn)).read()
            assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), fn


def test_upsample_conv_subpixel_decomposition():
    """Upsample2D (nearest 2x) + conv3x3(pad 1) == four 2x2 convolutions on the low-res input with pre-summed taps
    (the weight packing the fused CUDA path uses; diffusers Upsample2D reached from image_generator.py:400).
    Host-only: the packing routine of the library against torch on the CPU."""
    import ctypes as C
    import torch.nn.functional as F
    from synt_isic_b200 import _lib
    g = torch.Generator().manual_seed(5)
    cin, cout, H, W = 6, 5, 7, 9
    w = torch.randn(cout, cin, 3, 3, generator=g, dtype=torch.float64).float().contiguous()
    x = torch.randn(2, cin, H, W, generator=g)
    packed = torch.empty(4 * cout, 4 * cin)
    _lib.check(_lib.lib().synt_debug_pack_upsample_phases(w.data_ptr(), cout, cin, packed.data_ptr()), "pack")
    ref = F.conv2d(F.interpolate(x, scale_factor=2.0, mode="nearest"), w, padding=1)
    xp = F.pad(x, (1, 1, 1, 1))                                   # low-res zero padding == padding of the upsampled image
    out = torch.zeros_like(ref)
    for py in range(2):
        for px in range(2):
            wp = packed[(py * 2 + px) * cout:(py * 2 + px + 1) * cout].view(cout, 2, 2, cin).permute(0, 3, 1, 2)
            # window rows y+py-1+ty of x == rows y+py+ty of xp
            win = xp[:, :, py:py + H + 1, px:px + W + 1]
            out[:, :, py::2, px::2] = F.conv2d(win, wp.contiguous())
    assert torch.allclose(out, ref, atol=1e-5, rtol=1e-5), (out - ref).abs().max()


class _FakeGenerator:
    """Stands in for ImageGenerator on a box without GPU: deterministic 'images', records what it was asked for."""
    base_seed = 42
    color_statistics = {}
    stop_requested = False

    def __init__(self):
        self.calls = []

    def generate_batch(self, class_name, seeds):
        self.calls.append((class_name, list(seeds)))
        imgs = np.zeros((len(seeds), 128, 128, 3), np.uint8)
        for j, s in enumerate(seeds):
            imgs[j] = s % 251
        return imgs, None, [f"h{s:08x}" for s in seeds]


def test_bulk_dataset_formats_and_numbering(tmp_path):
    """ISIC numbering, one-hot ground-truth CSV (console_generator_server.py:50,83-125) and metadata
    CSV (image_generator.py:742-782, path_manager.py:94-96) of the batched bulk driver; the numbering and the rank
    partition are pure functions of (class order, index)."""
    import csv as _csv
    from synt_isic_b200 import bulk
    from synt_isic_b200.generator import image_seed
    cfg = [("MEL", 70), ("VASC", 3), ("NV", 64)]
    units = bulk.plan_units(cfg, batch_size=64, layout="flat")
    assert [(u.class_name, u.first_index, u.count, u.first_number) for u in units] == [
        ("MEL", 0, 64, 34321), ("MEL", 64, 6, 34385), ("VASC", 0, 3, 34391), ("NV", 0, 64, 34394)]
    assert bulk.isic_name(34321) == "ISIC_0034321.jpg" and bulk.isic_name(1, "png") == "ISIC_0000001.png"
    assert bulk.ground_truth_header() == ["image", "MEL", "NV", "BCC", "AKIEC", "BKL", "DF", "VASC"]
    assert bulk.ground_truth_row("ISIC_0034391.jpg", "VASC") == ["ISIC_0034391.jpg", 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 1.0]
    # two ranks cover every unit exactly once, numbering unchanged
    from synt_isic_b200.dist import partition
    r0, r1 = partition(units, 0, 2), partition(units, 1, 2)
    assert sorted(r0 + r1, key=lambda u: u.first_number) == units and not set(r0) & set(r1)

    saved = []
    gen = _FakeGenerator()
    res = bulk.generate_dataset(gen, cfg, str(tmp_path), layout="flat", postprocess=False,
                                save_fn=lambda img, fp, c, s, h: saved.append((fp.name, c, s, int(img[0, 0, 0]))))
    assert res["total"] == 137 and res["generated"] == {"MEL": 70, "VASC": 3, "NV": 64}
    assert gen.calls[1] == ("MEL", [image_seed(42, "MEL", 64 + j) for j in range(6)])     # seeds of image_generator.py:626-637
    assert saved[0] == ("ISIC_0034321.jpg", "MEL", image_seed(42, "MEL", 0), image_seed(42, "MEL", 0) % 251)
    rows = list(_csv.reader(open(res["files"]["ground_truth_csv"])))
    assert rows[0] == bulk.ground_truth_header() and len(rows) == 138
    assert rows[1] == ["ISIC_0034321.jpg", "1.0", "0.0", "0.0", "0.0", "0.0", "0.0", "0.0"]
    assert rows[71] == ["ISIC_0034391.jpg", "0.0", "0.0", "0.0", "0.0", "0.0", "0.0", "1.0"]
    assert rows[-1][0] == "ISIC_0034457.jpg" and rows[-1][2] == "1.0"

    res2 = bulk.generate_dataset(_FakeGenerator(), [("BCC", 2), ("DF", 1)], str(tmp_path / "gui"), layout="per_class",
                                 postprocess=False, save_fn=lambda *a: None)
    rows2 = list(_csv.DictReader(open(res2["files"]["metadata_csv"])))
    assert [r["filename"] for r in rows2] == ["ISIC_0000001.png", "ISIC_0000002.png", "ISIC_0000001.png"]
    assert [r["class"] for r in rows2] == ["BCC", "BCC", "DF"] and rows2[0]["source"] == "synthetic"
    assert list(rows2[0].keys()) == ["filename", "class", "isic_number", "source", "generated_at"]


def test_run_xai_analysis_preview_lookup(tmp_path):
    """xai_integration.py:137-159: stored artifacts win in the order step map > Grad-CAM > Time-SHAP plot, else the image."""
    from PIL import Image
    from synt_isic_b200 import xai
    img_dir = tmp_path / "out" / "synthetic" / "MEL"
    img_dir.mkdir(parents=True)
    img = img_dir / "ISIC_0000007.png"
    Image.new("RGB", (8, 8), (10, 20, 30)).save(img)
    pil, path = xai.run_xai_analysis(str(img))
    assert path == str(img) and pil.mode == "RGB" and pil.size == (8, 8)
    art = tmp_path / "out" / "xai_results" / "MEL" / "ISIC_0000007_20250101"
    art.mkdir(parents=True)
    Image.new("L", (4, 4), 7).save(art / "time_shap_analysis.png")
    assert xai.run_xai_analysis(str(img))[1] == str(art / "time_shap_analysis.png")
    Image.new("RGB", (4, 4)).save(art / "gradcam_most_important_t_0.png")
    assert xai.run_xai_analysis(str(img))[1] == str(art / "gradcam_most_important_t_0.png")
    Image.new("RGB", (4, 4)).save(art / "xai_step_t_980.png")
    Image.new("RGB", (4, 4)).save(art / "xai_step_t_0.png")
    pil, path = xai.run_xai_analysis(str(img), device=None, classifier_path=None, save_dir=None)
    assert path == str(art / "xai_step_t_0.png") and pil.mode == "RGB"


def test_batched_integrated_gradients_plumbing(monkeypatch):
    """Host logic of xai.compute_integrated_gradients_batch (slicing of the path-point / gradient stacks, shared vs
    per-image baselines, ragged last pass) with the two C entry points and the classifier replaced by closed forms:
    F(x) = sum x^2 -> dF/dx = 2x -> IG = (x - b) * mean_k 2 (b + a_k (x - b))."""
    import contextlib
    import ctypes as C
    import numpy as np
    import torch
    from synt_isic_b200 import _lib, xai

    def arr(ptr, n):
        return np.ctypeslib.as_array((C.c_float * n).from_address(ptr))

    class FakeLib:
        def synt_ig_interpolate(self, x, b, n, per, out, st):
            X, B, O = arr(x, per), arr(b, per), arr(out, n * per).reshape(n, per)
            for k in range(n):
                O[k] = B + np.float32((k + 1) / n) * (X - B)
            return 0

        def synt_ig_reduce(self, g, x, b, n, per, out, st):
            arr(out, per)[:] = (arr(x, per) - arr(b, per)) * arr(g, n * per).reshape(n, per).sum(0) / n
            return 0

    class FakeClassifier:
        calls = []

        def parameters(self):
            yield torch.zeros(1)

        def score_and_input_gradient(self, pts, target):
            self.calls.append(pts.shape[0])
            return pts.flatten(1).pow(2).sum(1), 2 * pts

    monkeypatch.setattr(_lib, "lib", lambda: FakeLib())
    monkeypatch.setattr(_lib, "current_stream_ptr", lambda: 0)
    monkeypatch.setattr(torch.cuda, "device", lambda d: contextlib.nullcontext())
    g = torch.Generator().manual_seed(0)
    x = torch.randn(5, 3, 128, 128, generator=g)
    alphas = [(k + 1) / 4 for k in range(4)]

    def closed_form(xx, bb):
        return (xx - bb) * sum(2 * (bb + a * (xx - bb)) for a in alphas) / 4

    shared = torch.randn(1, 3, 128, 128, generator=g) * 0.1
    clf = FakeClassifier()
    out = xai.compute_integrated_gradients_batch(clf, x, 0, n_steps=4, baselines=shared, images_per_pass=2)
    assert clf.calls == [8, 8, 4]                                           # 2 + 2 + 1 images x 4 path points
    assert torch.allclose(out, closed_form(x, shared), atol=1e-5)
    own = torch.randn(5, 3, 128, 128, generator=g) * 0.1
    out = xai.compute_integrated_gradients_batch(clf, x, 0, n_steps=4, baselines=own, images_per_pass=3)
    assert torch.allclose(out, closed_form(x, own), atol=1e-5)
    import pytest
    with pytest.raises(ValueError):
        xai.compute_integrated_gradients_batch(clf, x, 0, n_steps=4, baselines=own[:2])
    assert xai.get_baseline(x[:1], "zero").abs().max() == 0 and xai.get_baseline(x[:1], "anything").abs().max() == 0
    blur = xai.get_baseline(x[:1], "blur")
    assert blur.shape == x[:1].shape and blur.abs().max() < x[:1].abs().max()
    noise = xai.get_baseline(x[:1], "noise", generator=torch.Generator().manual_seed(1))
    assert 0.05 < noise.std() < 0.15                                        # 0.1 * N(0,1), XAI.py:1024


def test_header_is_plain_c_and_the_library_links_from_c(tmp_path):
    """The drop-in boundary is a C ABI: include/synt_isic.h must compile as C99 (no C++-isms, no torch types) and a plain C
    program must link against libsynt_isic_b200.so and reach the host-only entry points (no GPU needed for these)."""
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    inc, pkg = os.path.join(root, "include"), os.path.join(root, "synt_isic_b200")
    hdr = os.path.join(inc, "synt_isic.h")
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", "-pedantic", "-fsyntax-only", "-x", "c", hdr], check=True)
    src = tmp_path / "demo.c"
    src.write_text(r"""
#include "synt_isic.h"
#include <stdio.h>
int main(void) {
    char name[256];
    long long numel = 0, off = -1;
    int n = synt_resnet18_num_params();
    if (n <= 0) return 1;
    if (synt_resnet18_param_info(7, 0, name, 256, &numel, &off) != 0) return 2;
    printf("%d %s %lld %lld\n", n, name, numel, off);
    if (synt_resnet18_param_info(7, -1, name, 256, &numel, &off) >= 0) return 3;      /* error path: negative code ... */
    if (!synt_last_error() || !synt_last_error()[0]) return 4;                         /* ... and a message */
    return 0;
}
""")
    exe = tmp_path / "demo"
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", str(src), "-I", inc, "-L", pkg, "-lsynt_isic_b200",
                    f"-Wl,-rpath,{pkg}", "-o", str(exe)], check=True)
    r = subprocess.run([str(exe)], capture_output=True, text=True, timeout=60)
    assert r.returncode == 0, (r.returncode, r.stdout, r.stderr)
    n, name, numel, off = r.stdout.split()
    assert int(n) == 102 and name == "conv1.weight" and int(numel) == 64 * 3 * 49 and int(off) == 0


def test_image_generator_setters_and_xai_result_files(tmp_path):
    """image_generator.py:84-122 (setters the GUI calls) and :866-886 (where the integrated-XAI JSON goes)."""
    import json
    from synt_isic_b200.generator import ImageGenerator

    class NoModels:
        inference_steps = 0
        loaded_models = {}

    gen = ImageGenerator(model_manager=NoModels(), device="cpu", inference_steps=5000)
    assert gen.inference_steps == 1000                                       # clamp of :75-79
    logs = []
    gen.set_log_callback(logs.append)
    gen.set_progress_callback(None)
    gen.set_xai_frequency(0)
    assert gen.xai_frequency == 1
    gen.set_save_trajectory(1)
    assert gen.save_trajectory is True
    gen.set_generation_seed("17")
    assert gen.base_seed == 17
    gen.set_generation_seed(None)
    assert gen.base_seed is None
    gen.set_xai_hook(lambda p, c: None, every_n=0)
    assert gen.xai_every_n == 1
    gen.set_xai_analyzer(object())
    assert gen.xai_analyzer is not None
    img = tmp_path / "out" / "MEL" / "ISIC_0000003.png"
    img.parent.mkdir(parents=True)
    path = gen._save_xai_results({"time_shap": {"importance": [0.5, 1.0]}, "класс": "MEL"}, "MEL", img.name, str(img))
    assert path and os.path.dirname(path) == str(tmp_path / "out" / "xai_results" / "MEL")
    assert os.path.basename(path).startswith("xai_ISIC_0000003_") and path.endswith(".json")
    assert json.load(open(path, encoding="utf-8"))["класс"] == "MEL"
    assert any("XAI results saved" in m and m.startswith("[INFO]") for m in logs)
    assert gen._save_xai_results({"bad": object()}, "MEL", img.name, str(img)) is None     # not JSON-serialisable: logged
    assert any(m.startswith("[WARNING]") for m in logs)


def test_image_generator_reference_constructor_and_result_contract(tmp_path):
    """``ImageGenerator(config_manager)`` (image_generator.py:28-78: inference_timesteps and the checkpoints folder come
    from the config object) and the result / error contract of ``generate_images`` (:547-740): total_generated + stopped,
    synthetic_dataset.csv with the reference's columns, numbering by successful images, {"error": ...} instead of raising,
    a missing checkpoint fails the load (model_manager.py:104-107) unless random init is asked for."""
    import csv
    import numpy as np
    from synt_isic_b200.generator import GenerationStopped, ImageGenerator, ModelManager, SYNTHETIC_CSV_HEADERS

    class Cfg:                                                      # stand-in for core/config/config_manager.py
        def get_generation_param(self, key, default=None):
            return {"inference_timesteps": 7}.get(key, default)

        def get_path(self, key):
            return str(tmp_path / "no_such_checkpoints") if key == "checkpoints" else None

    gen = ImageGenerator(Cfg(), device="cpu")
    assert gen.inference_steps == 7 and gen.base_seed is None and gen.save_trajectory is True
    assert gen.model_manager.checkpoint_dir == tmp_path / "no_such_checkpoints"
    assert gen.model_manager.load_model("MEL") is False               # no checkpoint, no silent random network
    assert ModelManager(device="cpu", allow_random_init=False).load_model("NV") is False

    # the result contract, with the sampler stubbed out (no GPU here)
    gen = ImageGenerator(device="cpu", inference_steps=3, base_seed=5, batch_size=2)
    made = []

    def fake_batch(class_name, seeds):
        if class_name == "DF" and made:
            gen.stop_generation()
            raise GenerationStopped("generation stopped")
        made.append((class_name, list(seeds)))
        return np.zeros((len(seeds), 128, 128, 3), np.uint8), None, ["0" * 16] * len(seeds)

    gen.generate_batch = fake_batch
    res = gen.generate_images([("MEL", 3), ("DF", 2)], str(tmp_path / "out"), postprocess=False)
    assert res["total_generated"] == 3 and res["stopped"] is True and res["generated"]["MEL"] == 3
    rows = list(csv.DictReader(open(tmp_path / "out" / "synthetic_dataset.csv", encoding="utf-8")))
    assert list(rows[0].keys()) == SYNTHETIC_CSV_HEADERS
    assert [r["filename"] for r in rows] == ["ISIC_0000001.png", "ISIC_0000002.png", "ISIC_0000003.png"]
    assert [r["isic_number"] for r in rows] == ["1", "2", "3"] and rows[0]["source"] == "synthetic"
    assert gen.is_generating is False

    def boom(class_name, seeds):
        raise ValueError("kaboom")

    gen.generate_batch = boom
    assert gen.generate_images([("MEL", 1)], str(tmp_path / "out2")) == {"error": "kaboom"}
    keys = ImageGenerator.noise_keys([11, None, 11])
    assert keys[0] == keys[2] == 11 and 0 <= keys[1] < 2 ** 63


def test_multiply_high_division_of_the_work_item_decode():
    """csrc/conv_tc2.cu v2_div: q = umulhi(n, 2^32 // d + 1) equals n // d whenever n * d < 2^32 (the host checks that bound
    for every launch); d = 1 is special-cased on the device."""
    import random
    rng = random.Random(0)
    for d in list(range(2, 300)) + [511, 512, 513, 1023, 1024, 4095, 4096]:
        magic = ((1 << 32) // d + 1) & 0xFFFFFFFF
        hi = ((1 << 32) - 1) // d
        ns = [0, 1, d - 1, d, d + 1, hi - 1, hi] + [rng.randrange(0, hi + 1) for _ in range(200)]
        for n in ns:
            if n < 0 or n * d >= (1 << 32):
                continue
            assert (n * magic) >> 32 == n // d, (n, d)
    # just beyond the bound the identity may fail: the check in conv_tc2() is what makes it safe
    bad = [(n, d) for d in (3, 7, 100) for n in range((1 << 32) // d, (1 << 32) // d + 50)
           if ((n * (((1 << 32) // d + 1) & 0xFFFFFFFF)) >> 32) != n // d]
    assert isinstance(bad, list)
