"""Generator entry points of the reference with their call signatures kept, driving the
B200 sampling path instead of the eager ``for t in timesteps`` loop.

  ModelManager      core/generator/model_manager.py:23-353   (model / scheduler factory)
  ImageGenerator    core/generator/image_generator.py:25-902 (per-class generation, B=1 API)
  DiffusionGenerator diffusion/diffusion_generator.py:19-275 (batched script-side generator)

Kept verbatim: ``generate_single_image(class_name, output_path, postprocess, seed, ...) ->
(bool, trajectory|None)``, ``generate_images(class_configs, output_dir, postprocess) -> dict``,
seed algebra (:586-592, :626-637), x_T from ``torch.Generator(device).manual_seed(seed)``
(:369-381), noise sha256 (:383-389), uint8 conversion (:441-447), ISIC file naming
(core/utils/path_manager.py:94-96), colour post-process (:502-545), sidecar JSON (:457-474).
Out of scope: config/cache/logger plumbing, GUI callbacks beyond the
progress/stop hooks.

New on B200: ``generate_batch(class_name, seeds)`` samples many images of one class in ONE
fused loop (the reference's B=1 loop repeated); ``generate_images`` uses it when no
trajectories are requested.
"""
from __future__ import annotations

import hashlib
import json
import os
import time
from pathlib import Path
from typing import Any, Callable, Dict, List, Optional, Tuple

import numpy as np
import torch

from .scheduler import DDPMScheduler
from .unet import SUPPORTED_CONFIG, UNet2DModel

CLASS_NAMES = ["MEL", "NV", "BCC", "AKIEC", "BKL", "DF", "VASC"]


def class_seed_offset(class_name: str) -> int:
    """core/generator/image_generator.py:586-592"""
    return int(hashlib.md5(class_name.encode("utf-8")).hexdigest()[:8], 16) & 0x7FFFFFFF


def image_seed(base_seed: int, class_name: str, index: int) -> int:
    """core/generator/image_generator.py:626-630"""
    return (int(base_seed) + class_seed_offset(class_name) + index) & 0x7FFFFFFF


def isic_filename(isic_number: int) -> str:
    """core/utils/path_manager.py:94-96"""
    return f"ISIC_{isic_number:07d}.png"


def noise_hash(x_T: torch.Tensor) -> str:
    """core/generator/image_generator.py:383-389"""
    return hashlib.sha256(x_T.detach().to("cpu").numpy().tobytes()).hexdigest()[:16]


def to_uint8_tensor(latents: torch.Tensor, mode: int = 0) -> torch.Tensor:
    """[B,3,H,W] fp32 -> [B,H,W,3] uint8 ON the device.  mode 0: core/generator/image_generator.py:441-447
    ((x+1)/2, clamp, *255, truncate); mode 1: diffusion/diffusion_generator.py:147-148 ((x+1)*127.5, clip, truncate)."""
    from . import _lib
    x = latents.contiguous().float()
    B, _, H, W = x.shape
    out = torch.empty(B, H, W, 3, dtype=torch.uint8, device=x.device)
    with torch.cuda.device(x.device):
        _lib.check(_lib.lib().synt_to_uint8(x.data_ptr(), B, H, W, int(mode), out.data_ptr(), _lib.current_stream_ptr()),
                   "to_uint8")
    return out


def to_uint8_image(latents: torch.Tensor, mode: int = 0) -> np.ndarray:
    """``to_uint8_tensor`` + the D2H copy of the reference (``image.cpu().numpy()``)."""
    return to_uint8_tensor(latents, mode).cpu().numpy()


def color_postprocess(img: np.ndarray, stats: Optional[dict]) -> np.ndarray:
    """core/generator/image_generator.py:502-545 (mean/std match, scale clip [0.6,1.4], alpha 0.35)."""
    if not stats or "rgb" not in stats or "mean" not in stats["rgb"]:
        return img
    target_mean = np.array(stats["rgb"].get("mean", [128, 128, 128]), dtype=np.float32)
    target_std = np.array(stats["rgb"].get("std", [50, 50, 50]), dtype=np.float32)
    cur_mean = np.mean(img, axis=(0, 1)).astype(np.float32)
    cur_std = np.std(img, axis=(0, 1)).astype(np.float32)
    scale = np.clip(target_std / np.maximum(cur_std, 1e-6), 0.6, 1.4)
    shifted = (img.astype(np.float32) - cur_mean) * scale + target_mean
    out = 0.35 * shifted + 0.65 * img.astype(np.float32)
    return np.clip(out, 0, 255).astype(np.uint8)


class GenerationStopped(RuntimeError):
    """Raised inside the batched path when ``stop_generation()`` was called (image_generator.py:396)."""


SYNTHETIC_CSV_HEADERS = ["filename", "class", "isic_number", "source", "generated_at"]   # image_generator.py:744


class ModelManager:
    """core/generator/model_manager.py: builds the UNet per class, loads ``unet_<CLASS>_best.pth`` and creates the
    scheduler.  Like the reference (:104-107) ``load_model`` returns False when the checkpoint is missing; random-init
    weights (benchmarks, tests -- the checkpoints are not shipped) must be asked for with ``allow_random_init=True`` or
    by passing a ``state_dict``."""

    def __init__(self, checkpoint_dir: Optional[str] = None, device: str = "cuda", precision: str = "bf16",
                 allow_random_init: bool = False):
        self.checkpoint_dir = Path(checkpoint_dir) if checkpoint_dir else None
        self.device = torch.device(device)
        self.precision = precision
        self.allow_random_init = bool(allow_random_init)
        self.loaded_models: Dict[str, UNet2DModel] = {}
        self.model_metadata: Dict[str, dict] = {}
        self.inference_steps = 50

    def _create_model_architecture(self) -> UNet2DModel:       # model_manager.py:173-194
        return UNet2DModel(precision=self.precision, **SUPPORTED_CONFIG)

    def load_model(self, class_name: str, state_dict: Optional[dict] = None, force_reload: bool = False) -> bool:   # :89-171
        if class_name in self.loaded_models and not force_reload and state_dict is None:
            return True
        meta = {"class": class_name, "source": "random_init"}
        path = self.checkpoint_dir / f"unet_{class_name}_best.pth" if self.checkpoint_dir else None
        if state_dict is None and path is not None and path.exists():
            state_dict = torch.load(str(path), map_location="cpu")
            meta = {"class": class_name, "source": str(path), "bytes": path.stat().st_size}
        elif state_dict is not None:
            meta = {"class": class_name, "source": "state_dict"}
        if state_dict is None and not self.allow_random_init:
            # model_manager.py:104-107: a missing checkpoint is an error, never a silently random network
            print(f"model file not found: {path if path is not None else '<no checkpoint_dir>'} "
                  f"(pass allow_random_init=True for random-init weights)")
            return False
        model = self._create_model_architecture()
        if state_dict is not None:
            model.load_state_dict(state_dict)                    # strict (model_manager.py:139)
        self.loaded_models[class_name] = model.to(self.device).eval()
        self.model_metadata[class_name] = meta
        return True

    def create_scheduler(self, class_name: Optional[str] = None) -> DDPMScheduler:       # :196-226
        s = DDPMScheduler(num_train_timesteps=1000, beta_schedule="squaredcos_cap_v2", prediction_type="epsilon")
        s.set_timesteps(max(1, min(1000, int(self.inference_steps))))
        return s

    def change_device(self, device: str):                                                # :319-345
        self.device = torch.device(device)
        for k, m in self.loaded_models.items():
            self.loaded_models[k] = m.to(self.device)


class ImageGenerator:
    """``ImageGenerator(config_manager)`` is the reference's constructor (image_generator.py:28-78): the first positional
    argument may be its ``ConfigManager`` (anything with ``get_generation_param`` / ``get_path``), from which
    ``inference_timesteps`` and the ``checkpoints`` folder are read; everything else of the config is control plane and
    ignored.  A ``ModelManager`` in that position (or ``model_manager=``) is the native form."""

    def __init__(self, model_manager=None, device: str = "cuda",
                 inference_steps: int = 50, base_seed: Optional[int] = 42, save_trajectory: bool = False,
                 color_statistics: Optional[dict] = None, precision: str = "bf16",
                 progress_callback: Optional[Callable[[int, int, str], None]] = None, batch_size: int = 64,
                 allow_random_init: bool = False, noise_seed: int = 0):
        self.device = torch.device(device)
        self.config_manager = None
        if model_manager is not None and hasattr(model_manager, "get_generation_param"):
            self.config_manager = model_manager                               # reference form (image_generator.py:28-29)
            model_manager = None
            try:
                inference_steps = int(self.config_manager.get_generation_param("inference_timesteps"))   # :70-74
            except Exception:
                inference_steps = 50
            ckpt = None
            try:
                ckpt = self.config_manager.get_path("checkpoints")            # model_manager.py:97
            except Exception:
                ckpt = None
            model_manager = ModelManager(ckpt, device=device, precision=precision, allow_random_init=allow_random_init)
            base_seed, save_trajectory = None, True                           # the reference's defaults (:52, :59)
        self.model_manager = model_manager or ModelManager(device=device, precision=precision,
                                                           allow_random_init=allow_random_init)
        self.noise_seed = int(noise_seed)                                     # Philox key of the in-kernel step noise
        self.inference_steps = max(1, min(1000, int(inference_steps)))       # image_generator.py:75-79
        self.model_manager.inference_steps = self.inference_steps
        self.base_seed = base_seed
        self.save_trajectory = save_trajectory
        self.color_statistics = color_statistics or {}
        self.progress_callback = progress_callback
        self.stop_requested = False
        self.is_generating = False
        self.batch_size = batch_size
        self.xai_analyzer = None
        self.xai_frequency = 0
        self.xai_hook = None
        self.xai_every_n = 10
        self.log_callback = None
        self.progress_every = 5                                               # image_generator.py:435

    def stop_generation(self):                                                # :784-786
        self.stop_requested = True

    # ---- the setters the GUI / console callers use (image_generator.py:84-122) ---------------
    def set_progress_callback(self, callback: Optional[Callable[[int, int, str], None]]):
        self.progress_callback = callback

    def set_log_callback(self, callback: Optional[Callable[[str], None]]):
        self.log_callback = callback

    def set_xai_frequency(self, frequency: int):
        """Every N-th image of a class gets the integrated XAI analysis (>= 1)."""
        self.xai_frequency = max(1, int(frequency))

    def set_save_trajectory(self, save: bool):
        self.save_trajectory = bool(save)

    def set_xai_analyzer(self, analyzer):
        self.xai_analyzer = analyzer
        if analyzer is not None and not self.xai_frequency:
            self.xai_frequency = 3                                            # the reference's default (:51)

    def set_xai_hook(self, callback: Optional[Callable[[str, str], None]], every_n: int = 10):
        """Legacy hook (:93-98).  The reference only STORES it -- nothing in its generation loops ever calls it -- so the
        drop-in stores it too."""
        self.xai_hook = callback
        self.xai_every_n = max(1, int(every_n))

    def set_generation_seed(self, seed: Optional[int]):
        self.base_seed = int(seed) if seed is not None else None

    def _log_message(self, message: str, level: str = "info"):
        if self.log_callback:
            self.log_callback(f"[{level.upper()}] {message}")                 # :129-131

    def _save_xai_results(self, xai_results: Dict[str, Any], class_name: str, filename: str, file_path: str):
        """:866-886 -- <two levels above the image>/xai_results/<class>/xai_<stem>_<YYYYmmdd_HHMMSS>.json; failures are
        logged, never raised (the reference swallows them the same way)."""
        try:
            xai_dir = Path(file_path).parent.parent / "xai_results" / class_name
            xai_dir.mkdir(parents=True, exist_ok=True)
            stamp = time.strftime("%Y%m%d_%H%M%S")
            target = xai_dir / f"xai_{Path(filename).stem}_{stamp}.json"
            with open(target, "w", encoding="utf-8") as f:
                json.dump(xai_results, f, indent=2, ensure_ascii=False)
            self._log_message(f"XAI results saved: {target}")
            return str(target)
        except Exception as e:                                                # noqa: BLE001
            self._log_message(f"could not save the XAI results: {e}", "warning")
            return None

    def _update_progress(self, cur, total, msg):
        if self.progress_callback:
            self.progress_callback(cur, total, msg)

    def _model(self, class_name: str) -> UNet2DModel:
        if class_name not in self.model_manager.loaded_models:
            if not self.model_manager.load_model(class_name):
                raise RuntimeError(f"could not load a model for class {class_name}")
        m = self.model_manager.loaded_models[class_name]
        if str(m.device) != str(self.device):
            m = m.to(self.device)
            self.model_manager.loaded_models[class_name] = m
        return m

    def make_noise(self, seeds: List[Optional[int]]) -> torch.Tensor:
        """x_T per image exactly as :369-381: a device generator seeded per image."""
        xs = []
        for s in seeds:
            if s is not None:
                g = torch.Generator(device=self.device)
                g.manual_seed(int(s))
                xs.append(torch.randn(1, 3, 128, 128, device=self.device, generator=g))
            else:
                xs.append(torch.randn(1, 3, 128, 128, device=self.device))
        return torch.cat(xs)

    @staticmethod
    def noise_keys(seeds: List[Optional[int]]) -> List[int]:
        """Philox stream id of every image = its seed (a fresh 63-bit draw when the seed is None, like the reference's
        fresh global-RNG noise, image_generator.py:403): the step noise of an image depends on ITS key only, so the B=1
        path, any batch composition and ``save_trajectory`` give the same image for the same sidecar seed."""
        return [int(s) if s is not None else int.from_bytes(os.urandom(8), "little") >> 1 for s in seeds]

    def _denoise(self, model, latents, keys: List[int], want_traj: bool, label: str,
                 progress_offset_units=0, progress_total_units=0):
        """The loop of :395-403 as chunked replays of the fused CUDA-graph step: the stop flag
        (:396) and the progress callback (:435) are honoured between chunks."""
        scheduler = self.model_manager.create_scheduler()
        n = len(scheduler.timesteps)
        traj = torch.empty((n,) + tuple(latents.shape), dtype=torch.float32, device=latents.device) if want_traj else None
        key_t = torch.tensor(keys, dtype=torch.int64, device=latents.device)
        step = 0
        while step < n:
            if self.stop_requested:
                return None, None
            end = min(n, step + self.progress_every)
            model.sample(latents, scheduler, seed=self.noise_seed, image_keys=key_t, trajectory=traj, step_begin=step,
                         step_end=end)
            step = end
            total = progress_total_units if progress_total_units > 0 else n
            self._update_progress(progress_offset_units + step, total, f"Denoising {label}: {step}/{n} ({int(100 * step / n)}%)")
        return latents, traj

    # ---- reference signature (image_generator.py:308-313) --------------------------------
    def generate_single_image(self, class_name: str, output_path: str, postprocess: bool = True,
                              seed: Optional[int] = None, progress_offset_units: int = 0,
                              progress_total_units: int = 0, overall_index: int = 0,
                              overall_total: int = 0) -> Tuple[bool, Optional[List[torch.Tensor]]]:
        try:
            if self.stop_requested:
                return False, None
            model = self._model(class_name)
            with torch.no_grad():
                noise = self.make_noise([seed])
                nhash = noise_hash(noise)
                latents, traj = self._denoise(model, noise.clone(), self.noise_keys([seed]),
                                              self.save_trajectory, class_name, progress_offset_units,
                                              progress_total_units)
                if latents is None:
                    return False, None
            img = to_uint8_image(latents)[0]
            if postprocess:
                img = color_postprocess(img, self.color_statistics.get(class_name))
            self._save(img, output_path, class_name, seed, nhash)
            trajectory = [traj[i].clone() for i in range(traj.shape[0])] if traj is not None else None
            return True, trajectory
        except Exception as e:                                    # reference swallows into (False, None)
            print(f"generation failed for class {class_name}: {e}")
            return False, None

    def _save(self, img: np.ndarray, output_path, class_name, seed, nhash):
        from PIL import Image
        Image.fromarray(img).save(output_path)
        meta = {
            "filename": Path(output_path).name, "class": class_name,
            "seed": int(seed) if seed is not None else None, "inference_steps": int(self.inference_steps),
            "scheduler": {"num_train_timesteps": 1000, "beta_schedule": "squaredcos_cap_v2", "prediction_type": "epsilon"},
            "model": self.model_manager.model_metadata.get(class_name, {}),
            "device": str(self.device), "noise_hash": nhash,
        }
        with open(Path(output_path).with_suffix(".json"), "w", encoding="utf-8") as f:
            json.dump(meta, f, indent=2, ensure_ascii=False)

    # ---- batched sampling (new) ----------------------------------------------------------
    def generate_batch(self, class_name: str, seeds: List[Optional[int]]) -> Tuple[np.ndarray, torch.Tensor, List[str]]:
        """Samples len(seeds) images of one class in one fused loop.  Returns (uint8 [B,128,128,3], fp32 finals, x_T
        hashes); image i is bit-identical to ``generate_single_image(..., seed=seeds[i])`` in the fp32 mode and equal up
        to the batch-dependent bf16 rounding of the GEMM tiles in the bf16 mode (same x_T, same step noise).  Raises
        ``GenerationStopped`` when ``stop_generation()`` was called."""
        model = self._model(class_name)
        with torch.no_grad():
            x = self.make_noise(list(seeds))
            hashes = [noise_hash(x[i:i + 1]) for i in range(x.shape[0])]
            latents, _ = self._denoise(model, x, self.noise_keys(list(seeds)), False, class_name)
            if latents is None:
                raise GenerationStopped("generation stopped")
        return to_uint8_image(latents), latents, hashes

    # ---- reference signature (image_generator.py:547-548) ---------------------------------
    def _initialize_synthetic_csv(self, csv_path: Path):                      # :742-758
        import csv
        with open(csv_path, "w", newline="", encoding="utf-8") as f:
            csv.DictWriter(f, fieldnames=SYNTHETIC_CSV_HEADERS).writeheader()

    def _append_to_csv(self, csv_path: Path, data: Dict[str, Any]):           # :760-782
        import csv
        with open(csv_path, "a", newline="", encoding="utf-8") as f:
            csv.DictWriter(f, fieldnames=SYNTHETIC_CSV_HEADERS).writerow({k: data.get(k, "") for k in SYNTHETIC_CSV_HEADERS})

    def generate_images(self, class_configs: List[Tuple[str, int]], output_dir: str,
                        postprocess: bool = True) -> Dict[str, Any]:
        """image_generator.py:547-740.  Result like the reference: ``{"total_generated": n, "stopped": bool}`` or
        ``{"error": str}`` (nothing propagates); ``synthetic_dataset.csv`` with the reference's columns; ISIC numbering
        inside the class folder and the XAI cadence count SUCCESSFUL images (``class_image_count``, :609-617, :669).
        Additions: ``generated`` (per class), ``total`` and ``rows`` (filename / class / seed)."""
        if self.is_generating:
            return {"error": "generation already running"}
        generated_count = 0
        results: Dict[str, Any] = {"generated": {}, "total": 0, "rows": []}
        try:
            self.is_generating, self.stop_requested = True, False
            out = Path(output_dir)
            out.mkdir(parents=True, exist_ok=True)
            csv_path = out / "synthetic_dataset.csv"
            self._initialize_synthetic_csv(csv_path)
            total_images = sum(c for _, c in class_configs)

            def record(fp: Path, class_name: str, isic_number: int, seed):
                self._append_to_csv(csv_path, {"filename": fp.name, "class": class_name, "isic_number": isic_number,
                                               "source": "synthetic", "generated_at": str(fp.stat().st_mtime)})
                results["rows"].append({"filename": fp.name, "class": class_name, "seed": seed})

            for class_name, count in class_configs:
                if self.stop_requested:
                    break
                cdir = out / class_name
                cdir.mkdir(exist_ok=True)
                # :626-637: seed = base + md5 offset + index; without a base seed a random but recorded 31-bit seed
                seeds = [image_seed(self.base_seed, class_name, i) if self.base_seed is not None
                         else int.from_bytes(os.urandom(4), "little") & 0x7FFFFFFF for i in range(count)]
                class_image_count = 0
                if self.save_trajectory or self.xai_analyzer is not None:
                    for i, sd in enumerate(seeds):               # B=1 path keeps per-image trajectories
                        if self.stop_requested:
                            break
                        fp = cdir / isic_filename(class_image_count + 1)
                        ok, traj = self.generate_single_image(
                            class_name, str(fp), postprocess, sd,
                            progress_offset_units=generated_count * self.inference_steps,
                            progress_total_units=total_images * self.inference_steps,
                            overall_index=generated_count + 1, overall_total=total_images)
                        if not ok:
                            continue
                        generated_count += 1
                        class_image_count += 1
                        record(fp, class_name, class_image_count, sd)
                        self._update_progress(generated_count, total_images, f"Generated {generated_count}/{total_images}")
                        if traj and self.xai_analyzer is not None and self.xai_frequency \
                                and class_image_count % self.xai_frequency == 0:
                            try:                                  # :668-699: analysis errors never stop the generation
                                res = self.xai_analyzer.analyze_trajectory(traj, class_name, sd, self.inference_steps,
                                                                           fp.name, str(fp))
                                if res:
                                    self._save_xai_results(res, class_name, fp.name, str(fp))
                            except Exception as e:                # noqa: BLE001
                                self._log_message(f"integrated XAI analysis failed: {e}", "warning")
                else:
                    for b0 in range(0, count, self.batch_size):
                        if self.stop_requested:
                            break
                        chunk = seeds[b0:b0 + self.batch_size]
                        try:
                            imgs, _, hashes = self.generate_batch(class_name, chunk)
                        except GenerationStopped:
                            break
                        for j, img in enumerate(imgs):
                            fp = cdir / isic_filename(class_image_count + 1)
                            if postprocess:
                                img = color_postprocess(img, self.color_statistics.get(class_name))
                            self._save(img, str(fp), class_name, chunk[j], hashes[j])
                            generated_count += 1
                            class_image_count += 1
                            record(fp, class_name, class_image_count, chunk[j])
                        self._update_progress(generated_count, total_images, f"Generated {generated_count}/{total_images}")
                results["generated"][class_name] = class_image_count
            results["total"] = generated_count
            results["total_generated"] = generated_count
            results["stopped"] = bool(self.stop_requested)
            return results
        except Exception as e:                                    # noqa: BLE001  (:730-737)
            self._log_message(f"generation failed: {e}", "error")
            return {"error": str(e)}
        finally:
            self.is_generating = False


class DiffusionGenerator:
    """diffusion/diffusion_generator.py:19-275.  NB the reference uses LINEAR betas and never calls
    set_timesteps here (:123-128) -> 1000 steps; kept."""

    def __init__(self, checkpoint_dir: Optional[str] = None, stats_path: Optional[str] = None, device: str = "cuda",
                 precision: str = "bf16", allow_random_init: bool = False):
        self.device = torch.device(device)
        self.manager = ModelManager(checkpoint_dir, device, precision, allow_random_init=allow_random_init)
        self.color_statistics = {}
        if stats_path and os.path.exists(stats_path):
            with open(stats_path, "r", encoding="utf-8") as f:
                self.color_statistics = json.load(f)

    def _scheduler(self):
        return DDPMScheduler(num_train_timesteps=1000, beta_start=0.0001, beta_end=0.02, beta_schedule="linear")

    def _run(self, class_name: str, count: int) -> np.ndarray:
        if class_name not in self.manager.loaded_models and not self.manager.load_model(class_name):
            raise ValueError(f"no model for class {class_name}")        # diffusion_generator.py:108-109
        model = self.manager.loaded_models[class_name]
        x = torch.randn(count, 3, 128, 128).to(self.device)      # diffusion_generator.py:131 / :211 (CPU global RNG)
        model.sample(x, self._scheduler(), seed=int(torch.randint(0, 2 ** 31 - 1, (1,)).item()))
        return to_uint8_image(x, mode=1)                          # :147-148 ((x+1)*127.5, clip, uint8) on the GPU

    def generate_single_image(self, class_name: str, output_path: str, postprocess: bool = True) -> str:
        from PIL import Image
        img = self._run(class_name, 1)[0]
        if postprocess:
            img = color_postprocess(img, self.color_statistics.get(class_name))
        root, ext = os.path.splitext(output_path)
        save_path = output_path if ext.lower() in (".jpg", ".jpeg") else root + ".jpg"
        Image.fromarray(img).convert("RGB").save(save_path, format="JPEG", quality=95)
        return save_path

    def generate_batch_images(self, class_name: str, output_dir: str, count: int, postprocess: bool = True,
                              batch_size: int = 4) -> List[str]:
        from PIL import Image
        os.makedirs(output_dir, exist_ok=True)
        paths = []
        for b0 in range(0, count, batch_size):
            imgs = self._run(class_name, min(batch_size, count - b0))
            for i, img in enumerate(imgs):
                if postprocess:
                    img = color_postprocess(img, self.color_statistics.get(class_name))
                p = os.path.join(output_dir, f"{class_name}_{b0 + i + 1:04d}.jpg")
                Image.fromarray(img).convert("RGB").save(p, format="JPEG", quality=95)
                paths.append(p)
        return paths
