"""Build a VARIANT of libhoigen_b200.so with extra nvcc defines into hoigen_b200/_lib/variants/<name>/ (git-ignored, travels to
the GPU box) for same-box A/B runs:  python tools/build_variant.py NAME -DFOO=1 -DBAR=2 ;  then HOIGEN_B200_LIB=<path> python bench.py"""
import subprocess
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from hoigen_b200 import _build  # noqa: E402

name, defs = sys.argv[1], sys.argv[2:]
out_dir = _build.LIB_DIR / "variants" / name
(out_dir / "obj").mkdir(parents=True, exist_ok=True)
nvcc = _build._nvcc()
procs, objs = [], []
for src in _build._sources():
    obj = out_dir / "obj" / (src.stem + ".o")
    objs.append(str(obj))
    procs.append((src, subprocess.Popen([nvcc, *_build.NVCC_FLAGS, *defs, "-c", str(src), "-o", str(obj)],
                                        stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
for src, p in procs:
    out, _ = p.communicate()
    if p.returncode != 0:
        raise SystemExit(f"nvcc failed on {src.name}:\n{out}")
lib = out_dir / "libhoigen_b200.so"
subprocess.run([nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-cudart", "static", "-Xcompiler", "-fPIC",
                "-o", str(lib), *objs, "-ldl", "-lpthread", "-lrt"], check=True)
for o in objs:
    Path(o).unlink()
print(lib)
