"""Build libdroneyolo.so (the C-ABI CUDA library) in-tree with nvcc for sm_100a.

The library is plain CUDA C++ with an `extern "C"` surface (include/droneyolo.h); it does not link torch.
`python -m drone_yolo_b200.build` (or `__graft_entry__.build()`) cross-compiles without a GPU.
"""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

PKG = Path(__file__).resolve().parent
CSRC = PKG / "csrc"
OUT_DIR = PKG / "lib"
DEBUG = bool(os.environ.get("DY_CONV_DEBUG_BUILD"))           # knock-outs + in-kernel timeline (tools/bench_conv.py, trace_conv.py)
KNOCK = os.environ.get("DY_CONV_KNOCKOUT_BUILD", "")          # compile-time knock-out mask: release-speed "what bounds it" builds
EXTRA = os.environ.get("DY_EXTRA_NVCC_FLAGS", "")              # experiment builds: extra -D flags, library tagged DY_LIB_TAG
_TAG = "_dbg" if DEBUG else (f"_k{KNOCK}" if KNOCK else (("_" + os.environ.get("DY_LIB_TAG", "x")) if EXTRA else ""))
LIB = OUT_DIR / f"libdroneyolo{_TAG}.so"
OBJ_DIR = PKG / f"build{_TAG}"

NVCC_FLAGS = [
    "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
    "-Xcompiler", "-fPIC", "-Xptxas", "-v", "--expt-relaxed-constexpr",
] + (["-DDY_CONV_DEBUG"] if os.environ.get("DY_CONV_DEBUG_BUILD") else []) + (
    [f"-DDY_CONV_DBG_CONST={int(os.environ['DY_CONV_KNOCKOUT_BUILD'])}"] if os.environ.get("DY_CONV_KNOCKOUT_BUILD") else []) + EXTRA.split()


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and Path(cand).exists():
            return cand
    raise RuntimeError("nvcc not found; libdroneyolo.so cannot be built")


def _sources() -> list[Path]:
    return sorted(CSRC.glob("*.cu"))


def _digest() -> str:
    h = hashlib.sha256()
    for f in sorted(list(CSRC.glob("*")) + [PKG.parent / "include" / "droneyolo.h"]):
        h.update(f.name.encode())
        h.update(f.read_bytes())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def build(force: bool = False, verbose: bool = False) -> Path:
    """Compile every .cu under csrc/ and link lib/libdroneyolo.so. Skips when sources are unchanged."""
    OUT_DIR.mkdir(exist_ok=True)
    OBJ_DIR.mkdir(exist_ok=True)
    stamp = OUT_DIR / (LIB.stem + ".sha256")
    digest = _digest()
    if not force and LIB.exists() and stamp.exists() and stamp.read_text().strip() == digest:
        return LIB
    nvcc = _nvcc()
    srcs = _sources()

    def compile_one(src: Path) -> tuple[Path, str]:
        obj = OBJ_DIR / (src.stem + ".o")
        cmd = [nvcc, *NVCC_FLAGS, "-c", str(src), "-o", str(obj)]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src.name}:\n{r.stdout}\n{r.stderr}")
        return obj, r.stderr

    with ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        results = list(ex.map(compile_one, srcs))
    log = "\n".join(f"== {o.name}\n{msg}" for o, msg in results)
    (OBJ_DIR / "ptxas.log").write_text(log)
    if verbose:
        print(log)
    cmd = [nvcc, "-shared", "-o", str(LIB), *[str(o) for o, _ in results], "-gencode", "arch=compute_100a,code=sm_100a"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    stamp.write_text(digest)
    return LIB


if __name__ == "__main__":
    path = build(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(path)
