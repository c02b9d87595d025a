"""Builds ``libsynt_isic_b200.so`` in-tree with nvcc for sm_100a (no GPU needed to build).

    python -m synt_isic_b200.build [--force]

The shared library is a plain C-ABI library (include/synt_isic.h); it links the CUDA
runtime statically and resolves ``cuTensorMapEncodeTiled`` from the driver at run time, so
it has no link-time dependency on libcuda or on PyTorch.
"""
from __future__ import annotations

import hashlib
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

PKG = Path(__file__).resolve().parent
CSRC = PKG / "csrc"
BUILD = PKG / "_build"
LIB = PKG / "libsynt_isic_b200.so"
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
CFLAGS = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr"]
CFLAGS += os.environ.get("SYNT_EXTRA_NVCC_FLAGS", "").split()      # e.g. -DSYNT_ATT_TIMELINE_BUILD for tools/att_timeline.py
# SYNT_EXPERIMENTS=1: the measurement scaffolding (role knock-outs of conv_tc2, the halo-descriptor experiment kernel) is
# compiled in; the product build carries none of it
EXPERIMENTS = os.environ.get("SYNT_EXPERIMENTS", "0") == "1"
if EXPERIMENTS:
    CFLAGS.append("-DSYNT_EXPERIMENTS")
EXPERIMENT_SOURCES = {"conv_tc_halo.cu"}


def _sources():
    return sorted(p for p in CSRC.glob("*.cu") if EXPERIMENTS or p.name not in EXPERIMENT_SOURCES)


def _fingerprint() -> str:
    h = hashlib.sha256()
    for p in sorted(list(CSRC.glob("*")) + [PKG.parent / "include" / "synt_isic.h"]):
        if p.is_file():
            h.update(p.name.encode())
            h.update(p.read_bytes())
    h.update(" ".join(ARCH + CFLAGS).encode())
    return h.hexdigest()


def build(force: bool = False, verbose: bool = True) -> Path:
    BUILD.mkdir(exist_ok=True)
    stamp = BUILD / "fingerprint"
    fp = _fingerprint()
    if not force and LIB.exists() and stamp.exists() and stamp.read_text() == fp:
        return LIB
    srcs = _sources()

    def compile_one(src: Path):
        obj = BUILD / (src.stem + ".o")
        cmd = [NVCC, *ARCH, *CFLAGS, "-c", str(src), "-o", str(obj)]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src.name}:\n{r.stdout}\n{r.stderr}")
        return obj

    with ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        objs = list(ex.map(compile_one, srcs))
    cmd = [NVCC, *ARCH, "-shared", "-o", str(LIB), *map(str, objs)]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    stamp.write_text(fp)
    if verbose:
        print(f"[synt_isic_b200] built {LIB} from {len(srcs)} sources", file=sys.stderr)
    return LIB


if __name__ == "__main__":
    build(force="--force" in sys.argv)
