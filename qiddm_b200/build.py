"""Builds libqiddm_b200.so in-tree with nvcc for sm_100a (no JIT cache: the .so travels with the repo)."""
from __future__ import annotations

import os
import shutil
import subprocess
from pathlib import Path

PKG = Path(__file__).resolve().parent
ROOT = PKG.parent
CSRC = PKG / "csrc"
LIB_DIR = PKG / "lib"
LIB_PATH = LIB_DIR / "libqiddm_b200.so"
SOURCES = ["qiddm_gate.cu", "qiddm_gemm.cu", "qiddm_pca.cu", "qiddm_glue.cu", "qiddm_dm.cu", "qiddm_linear.cu", "qiddm_conv.cu", "qiddm_api.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC"]


def find_nvcc() -> str:
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found: libqiddm_b200.so cannot be built")
    return nvcc


def sources() -> list[Path]:
    return [CSRC / s for s in SOURCES if (CSRC / s).exists()]


def is_stale() -> bool:
    if not LIB_PATH.exists():
        return True
    t = LIB_PATH.stat().st_mtime
    deps = list(CSRC.glob("*.cu")) + list(CSRC.glob("*.h")) + list(CSRC.glob("*.cuh")) + [ROOT / "include" / "qiddm.h"]
    return any(d.stat().st_mtime > t for d in deps if d.exists())


def build(force: bool = False, verbose: bool = False) -> Path:
    if not force and not is_stale():
        return LIB_PATH
    LIB_DIR.mkdir(exist_ok=True)
    obj_dir = PKG / "build"
    obj_dir.mkdir(exist_ok=True)
    nvcc = find_nvcc()
    inc = [f"-I{ROOT / 'include'}", f"-I{CSRC}"]
    # one nvcc per translation unit, in parallel; then link
    procs, objs = [], []
    hdr_t = max(d.stat().st_mtime for d in [*CSRC.glob("*.h"), *CSRC.glob("*.cuh"), ROOT / "include" / "qiddm.h"] if d.exists())
    for src in sources():
        obj = obj_dir / (src.stem + ".o")
        objs.append(obj)
        if not force and obj.exists() and obj.stat().st_mtime > max(src.stat().st_mtime, hdr_t):
            continue          # this translation unit is up to date
        cmd = [nvcc, *NVCC_FLAGS, *inc, "-c", "-o", str(obj), str(src)]
        if verbose:
            print(" ".join(cmd))
        procs.append((cmd, subprocess.Popen(cmd)))
    for cmd, pr in procs:
        if pr.wait() != 0:
            raise subprocess.CalledProcessError(pr.returncode, cmd)
    link = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", str(LIB_PATH), *[str(o) for o in objs]]
    if verbose:
        print(" ".join(link))
    subprocess.run(link, check=True)
    return LIB_PATH


if __name__ == "__main__":
    print(build(force=True, verbose=True))
