"""Builds libotb200.so (hand-written sm_100a CUDA behind a C ABI) in-tree with nvcc.

The library is the product; there is no JIT and no fallback.  `build()` cross-compiles on a
GPU-less box (nvcc needs no device) and the resulting .so travels with the repo snapshot.
"""
from __future__ import annotations

import os
import shutil
import subprocess
from pathlib import Path

PKG_DIR = Path(__file__).resolve().parent
REPO_DIR = PKG_DIR.parent
CSRC = PKG_DIR / "csrc"
LIB_PATH = PKG_DIR / "libotb200.so"

SOURCES = ["api.cu", "optim.cu", "optim_factored.cu", "noise_loss.cu", "lora_gemm.cu", "lora_grad.cu", "dora.cu", "exchange.cu", "sr_round.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC",
    "--expt-relaxed-constexpr",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: libotb200.so cannot be built (there is no non-CUDA fallback)")


def _stale(target: Path, deps: list[Path]) -> bool:
    if not target.exists():
        return True
    t = target.stat().st_mtime
    return any(d.stat().st_mtime > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> Path:
    """Compile every .cu under csrc/ for sm_100a and link libotb200.so next to this file."""
    nvcc = _nvcc()
    obj_dir = PKG_DIR / "build"
    obj_dir.mkdir(exist_ok=True)
    headers = [CSRC / "common.cuh", CSRC / "tc_ptx.cuh", REPO_DIR / "include" / "otb200.h"]
    objs = []
    for src in SOURCES:
        s = CSRC / src
        o = obj_dir / (s.stem + ".o")
        if force or _stale(o, [s, *headers]):
            cmd = [nvcc, *NVCC_FLAGS, "-c", str(s), "-o", str(o)]
            if verbose:
                cmd[1:1] = ["-Xptxas", "-v"]
                print(" ".join(cmd))
            subprocess.run(cmd, check=True)
        objs.append(o)
    if force or _stale(LIB_PATH, objs):
        cmd = [nvcc, "-shared", "-o", str(LIB_PATH), *map(str, objs), "-Wno-deprecated-gpu-targets"]
        if verbose:
            print(" ".join(cmd))
        subprocess.run(cmd, check=True)
    return LIB_PATH


if __name__ == "__main__":
    import sys
    print(build(force="--force" in sys.argv, verbose=True))
