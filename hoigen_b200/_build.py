"""In-tree build of libhoigen_b200.so (sm_100a only) with nvcc.

The library has no torch / Python dependency: it is the C-ABI boundary declared in include/hoigen_b200.h.
nvcc cross-compiles without a GPU, so this runs in the CPU-only container; the resulting .so travels to the
GPU box with the repo snapshot (it is git-ignored, not gpurun-ignored).
"""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
from pathlib import Path

PKG_DIR = Path(__file__).resolve().parent
CSRC = PKG_DIR / "csrc"
LIB_DIR = PKG_DIR / "_lib"
LIB_PATH = LIB_DIR / "libhoigen_b200.so"
STAMP = LIB_DIR / "build.stamp"

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC",
    "-Xcompiler", "-fvisibility=hidden",
    "--expt-relaxed-constexpr",
    "-cudart", "static",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found (set NVCC=/path/to/nvcc)")


def _sources():
    return sorted(CSRC.glob("*.cu"))


def _fingerprint() -> str:
    h = hashlib.sha256()
    for p in sorted(list(CSRC.glob("*.cu")) + list(CSRC.glob("*.cuh")) + list(CSRC.glob("*.h"))
                    + [PKG_DIR.parent / "include" / "hoigen_b200.h"]):
        h.update(p.name.encode())
        h.update(p.read_bytes())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def build(force: bool = False, verbose: bool = False) -> Path:
    """Compile every .cu under csrc/ and link libhoigen_b200.so. Idempotent (content-hash stamp)."""
    LIB_DIR.mkdir(exist_ok=True)
    fp = _fingerprint()
    if not force and LIB_PATH.exists() and STAMP.exists() and STAMP.read_text().strip() == fp:
        return LIB_PATH
    nvcc = _nvcc()
    obj_dir = LIB_DIR / "obj"
    obj_dir.mkdir(exist_ok=True)
    procs = []
    objs = []
    for src in _sources():
        obj = obj_dir / (src.stem + ".o")
        objs.append(str(obj))
        cmd = [nvcc, *NVCC_FLAGS, "-c", str(src), "-o", str(obj)]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    failed = []
    for src, p in procs:
        out, _ = p.communicate()
        if verbose and out:
            print(out)
        if p.returncode != 0:
            failed.append(f"--- {src.name} ---\n{out}")
    if failed:
        raise RuntimeError("nvcc failed:\n" + "\n".join(failed))
    link = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-cudart", "static",
            "-Xcompiler", "-fPIC", "-o", str(LIB_PATH), *objs, "-ldl", "-lpthread", "-lrt"]
    r = subprocess.run(link, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n" + r.stdout)
    STAMP.write_text(fp)
    return LIB_PATH


if __name__ == "__main__":
    import sys
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
