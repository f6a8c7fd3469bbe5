"""Build ``libavld.so`` in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python -m amphibian_vae_latent_detector_b200.build [--force] [--verbose]

One ``-gencode arch=compute_100a,code=sm_100a`` image only: a plain ``-arch=sm_100a`` would also emit a
``compute_100`` PTX pass that rejects tcgen05.  ``rms.cu`` is compiled with ``-fmad=false`` because its
arithmetic must be bit-identical to numpy's (every multiply/add separately rounded).
"""
from __future__ import annotations

import argparse
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

HERE = Path(__file__).resolve().parent
CSRC = HERE / "csrc"
LIB = HERE / "libavld.so"
OBJ = HERE / "build"

SOURCES = ["ctx.cu", "rms.cu", "gemm3.cu", "fold3.cu", "dftf3.cu", "logmel.cu", "convh.cu", "encoder.cu", "radial.cu", "map.cu", "resample.cu", "comm.cu", "api.cu"]
EXTRA_FLAGS = {"rms.cu": ["-fmad=false"]}
BRINGUP_ONLY = ["conv1t.cu"]      # experiments that lost their A/B: built into libavld_bringup.so only
COMMON = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo",
          "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr"]


def _nvcc() -> str:
    cand = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not Path(cand).exists():
        raise RuntimeError("nvcc not found; libavld cannot be built")
    return cand


def _stale(target: Path, deps) -> bool:
    if not target.exists():
        return True
    t = target.stat().st_mtime
    return any(Path(d).stat().st_mtime > t for d in deps)


def build(force: bool = False, verbose: bool = False, bringup: bool = False) -> Path:
    """``bringup=True`` builds ``libavld_bringup.so`` with -DAVLD_BRINGUP (kernel probes and cycle counters driven by
    AVLD_DBG; used by tools/ only -- the product library has none of it compiled in)."""
    global LIB, OBJ
    nvcc = _nvcc()
    if bringup:
        LIB, OBJ = HERE / "libavld_bringup.so", HERE / "build_bringup"
    else:
        LIB, OBJ = HERE / "libavld.so", HERE / "build"
    OBJ.mkdir(exist_ok=True)
    headers = list(CSRC.glob("*.cuh")) + [HERE.parent / "include" / "avld.h", Path(__file__)]
    jobs = []
    sources = SOURCES + (BRINGUP_ONLY if bringup else [])
    for src in sources:
        obj = OBJ / (src + ".o")
        if force or _stale(obj, [CSRC / src] + headers):
            cmd = [nvcc, *COMMON, *EXTRA_FLAGS.get(src, []), *(["-DAVLD_BRINGUP"] if bringup else []), "-c", str(CSRC / src), "-o", str(obj)]
            if verbose:
                cmd.insert(1, "-Xptxas=-v")
            jobs.append(cmd)

    def run(cmd):
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + r.stdout + r.stderr)
        if verbose:
            print(r.stderr, file=sys.stderr)
        return 0

    if jobs:
        with ThreadPoolExecutor(max_workers=min(len(jobs), os.cpu_count() or 4)) as ex:
            list(ex.map(run, jobs))
    objs = [str(OBJ / (s + ".o")) for s in sources]
    if force or jobs or _stale(LIB, objs):
        cmd = [nvcc, "-shared", "-o", str(LIB), *objs, "-gencode", "arch=compute_100a,code=sm_100a",
               "-Xcompiler", "-fPIC"]   # static cudart: no loader-path dependency on the GPU box
        run(cmd)
    return LIB


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--force", action="store_true")
    ap.add_argument("--verbose", action="store_true")
    ap.add_argument("--bringup", action="store_true", help="build libavld_bringup.so (-DAVLD_BRINGUP) for tools/")
    a = ap.parse_args()
    print(build(force=a.force, verbose=a.verbose, bringup=a.bringup))
