"""Bulk synthetic-dataset driver and on-disk formats.

What the reference does (all paths relative to the reference repo):

* ``diffusion/console_generator_server.py:405-467`` -- ONE folder ``ISIC2018_Task3_synt`` with a global ISIC
  numbering that continues after the last number of the real data set (34320, ``:50``), file names
  ``ISIC_%07d.jpg`` (``:83-86``) and a one-hot ground-truth CSV ``ISIC2018_Task3_GroundTruth_synt.csv`` with the
  header ``image,MEL,NV,BCC,AKIEC,BKL,DF,VASC`` and ``1.0`` / ``0.0`` cells (``:88-125``); one model load and one
  B=1 sampling run per image.
* ``core/generator/image_generator.py:547-740`` -- one folder per class, numbering restarts at 1 inside each class
  folder (``:616-619``), ``ISIC_%07d.png`` (``core/utils/path_manager.py:94-96``), a metadata CSV with the columns
  ``filename,class,isic_number,source,generated_at`` (``:742-782``) and a JSON side-car per image (``:457-474``).

Here the same files come out of BATCHED sampling (64 images of one class per pass of the CUDA path, one model load per
class) and the (class, batch) units are spread round-robin over the ranks of a ``torch.distributed`` group: every rank
writes its own image files, rank 0 writes the CSV files after ONE ``gather_object`` of the rows (the only collective).
The numbering is a pure function of (class order, index), so it does not depend on the number of ranks.
"""
from __future__ import annotations

import csv
import time
from dataclasses import dataclass
from pathlib import Path
from typing import Callable, Dict, Iterable, List, Optional, Sequence, Tuple

import numpy as np

from .classifier import CLASS_NAMES
from .dist import partition

LAST_REAL_ISIC_NUMBER = 34320                       # console_generator_server.py:50
GROUND_TRUTH_CSV = "ISIC2018_Task3_GroundTruth_synt.csv"
SYNTHETIC_DIR = "ISIC2018_Task3_synt"
METADATA_CSV = "synthetic_metadata.csv"
METADATA_COLUMNS = ["filename", "class", "isic_number", "source", "generated_at"]   # image_generator.py:745


def isic_name(number: int, ext: str = "jpg") -> str:
    """``ISIC_0034321.jpg`` (console path) / ``.png`` (GUI path)."""
    return f"ISIC_{number:07d}.{ext}"


def ground_truth_header() -> List[str]:
    return ["image"] + list(CLASS_NAMES)                                             # console_generator_server.py:92-95


def ground_truth_row(image_name: str, class_name: str) -> list:
    """One-hot row exactly as ``_create_csv_row`` builds it (floats 1.0 / 0.0; unknown class -> all zeros)."""
    row = [image_name] + [0.0] * len(CLASS_NAMES)
    if class_name in CLASS_NAMES:
        row[CLASS_NAMES.index(class_name) + 1] = 1.0
    return row


@dataclass(frozen=True)
class Unit:
    """One batch of one class: the work unit that is assigned to a rank."""
    class_name: str
    first_index: int          # index of the first image inside its class (0-based)
    count: int
    first_number: int         # ISIC number of the first image


def plan_units(class_configs: Sequence[Tuple[str, int]], batch_size: int = 64, layout: str = "flat",
               start_number: int = LAST_REAL_ISIC_NUMBER) -> List[Unit]:
    """Splits ``[(class, count), ...]`` into (class, batch) units and fixes every image's ISIC number up front.
    layout "flat": one global numbering continuing after ``start_number`` in class_configs order (console path);
    layout "per_class": numbering restarts at 1 inside every class folder (GUI path)."""
    if layout not in ("flat", "per_class"):
        raise ValueError("layout must be 'flat' or 'per_class'")
    units, running = [], start_number
    for class_name, count in class_configs:
        if class_name not in CLASS_NAMES:
            raise ValueError(f"unknown ISIC class {class_name!r}")
        for b0 in range(0, int(count), batch_size):
            n = min(batch_size, int(count) - b0)
            first = running + b0 + 1 if layout == "flat" else b0 + 1
            units.append(Unit(class_name, b0, n, first))
        if layout == "flat":
            running += int(count)
    return units


def unit_rows(u: Unit, layout: str, ext: str, stamp: str) -> List[dict]:
    return [{"filename": isic_name(u.first_number + j, ext), "class": u.class_name, "isic_number": u.first_number + j,
             "source": "synthetic", "generated_at": stamp, "index_in_class": u.first_index + j} for j in range(u.count)]


def write_csvs(rows: Iterable[dict], root: Path, layout: str) -> Dict[str, str]:
    """Ground-truth CSV (flat layout, next to the image folder like the reference) or metadata CSV (per-class layout)."""
    rows = sorted(rows, key=lambda r: (CLASS_NAMES.index(r["class"]), r["isic_number"]) if layout == "per_class"
                  else r["isic_number"])
    out = {}
    if layout == "flat":
        p = root / GROUND_TRUTH_CSV
        with open(p, "w", newline="", encoding="utf-8") as f:
            w = csv.writer(f)
            w.writerow(ground_truth_header())
            for r in rows:
                w.writerow(ground_truth_row(r["filename"], r["class"]))
        out["ground_truth_csv"] = str(p)
    else:
        p = root / METADATA_CSV
        with open(p, "w", newline="", encoding="utf-8") as f:
            w = csv.DictWriter(f, fieldnames=METADATA_COLUMNS)
            w.writeheader()
            for r in rows:
                w.writerow({k: r[k] for k in METADATA_COLUMNS})
        out["metadata_csv"] = str(p)
    return out


def generate_dataset(generator, class_configs: Sequence[Tuple[str, int]], output_dir: str, layout: str = "flat",
                     postprocess: bool = True, batch_size: int = 64, group=None,
                     save_fn: Optional[Callable[[np.ndarray, Path, str, int, str], None]] = None) -> Dict[str, object]:
    """Generates the data set described by ``class_configs`` with ``generator`` (a ``synt_isic_b200.ImageGenerator``).

    Every rank of ``group`` (None = this process alone) must make the same call: the (class, batch) units are assigned
    round-robin, each rank samples its units with the batched CUDA path (``ImageGenerator.generate_batch``) and writes
    its own image files + JSON side-cars; rank 0 collects the rows with one ``gather_object`` and writes the CSV.
    Returns ``{"total", "generated": {class: n}, "rows", "files": {...}}`` on rank 0 (rows of all ranks) and the local
    part elsewhere."""
    import torch.distributed as dist
    from .generator import color_postprocess, image_seed

    rank, world = (dist.get_rank(group), dist.get_world_size(group)) if (group is not None and dist.is_initialized()) else (0, 1)
    root = Path(output_dir)
    ext = "jpg" if layout == "flat" else "png"
    img_root = root / SYNTHETIC_DIR if layout == "flat" else root
    img_root.mkdir(parents=True, exist_ok=True)
    units = plan_units(class_configs, batch_size, layout)
    mine = partition(units, rank, world)
    stamp = str(time.time())
    rows: List[dict] = []
    for u in mine:
        if getattr(generator, "stop_requested", False):
            break
        folder = img_root if layout == "flat" else img_root / u.class_name
        folder.mkdir(parents=True, exist_ok=True)
        seeds = [image_seed(generator.base_seed if generator.base_seed is not None else 42, u.class_name, u.first_index + j)
                 for j in range(u.count)]
        imgs, _, hashes = generator.generate_batch(u.class_name, seeds)
        for j, (img, r) in enumerate(zip(imgs, unit_rows(u, layout, ext, stamp))):
            if postprocess:
                img = color_postprocess(img, generator.color_statistics.get(u.class_name))
            fp = folder / r["filename"]
            if save_fn is not None:
                save_fn(img, fp, u.class_name, seeds[j], hashes[j])
            else:
                generator._save(img, str(fp), u.class_name, seeds[j], hashes[j])
            r["seed"] = seeds[j]
            rows.append(r)
    if world > 1:
        gathered = [None] * world if rank == 0 else None
        dist.gather_object(rows, gathered, dst=0, group=group)          # the one exchange of the bulk path
        if rank == 0:
            rows = [r for part in gathered for r in part]
    result: Dict[str, object] = {"rows": rows, "total": len(rows), "generated": {}}
    for r in rows:
        result["generated"][r["class"]] = result["generated"].get(r["class"], 0) + 1
    if rank == 0:
        result["files"] = write_csvs(rows, root, layout)
    return result
