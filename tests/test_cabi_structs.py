"""The ctypes mirrors in hoigen_b200/_cabi.py against include/hoigen_b200.h: a C program compiled with gcc from the header prints
sizeof / offsetof of every struct the ABI passes by pointer; every ctypes Structure must agree field by field.  (A silent
mismatch here would hand the kernels shifted pointers; GPU tests would catch it late and confusingly.)"""
import ctypes as C
import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent

MIRRORS = {
    "hoigen_gemm_params": "GemmParams",
    "hoigen_conv_op": "ConvOp",
    "hoigen_adapter_weights": "AdapterWeights",
    "hoigen_encoder_weights": "EncoderWeights",
    "hoigen_encoder_buffers": "EncoderBuffers",
    "hoigen_score_weights": "ScoreWeights",
    "hoigen_score_buffers": "ScoreBuffers",
    "hoigen_score_weights_fp32": "ScoreWeightsFp32",
    "hoigen_score_buffers_fp32": "ScoreBuffersFp32",
    "hoigen_folded_weights": "FoldedWeights",
}


def _c_name(field: str) -> str:
    return field[:-1] if field.endswith("_") else field      # `in_` mirrors `in` (a Python keyword)


def test_ctypes_mirrors_match_the_header(tmp_path):
    sys.path.insert(0, str(ROOT))
    from hoigen_b200 import _cabi
    lines = ["#include <stdio.h>", "#include <stddef.h>", '#include "hoigen_b200.h"', "int main(void) {"]
    expect = []
    for cname, pyname in MIRRORS.items():
        cls = getattr(_cabi, pyname)
        lines.append(f'  printf("%zu\\n", sizeof({cname}));')
        expect.append((f"sizeof({cname})", C.sizeof(cls)))
        for fname, _ftype in cls._fields_:
            lines.append(f'  printf("%zu\\n", offsetof({cname}, {_c_name(fname)}));')
            expect.append((f"offsetof({cname}, {_c_name(fname)})", getattr(cls, fname).offset))
    lines += ["  return 0;", "}"]
    src = tmp_path / "abi_layout.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "abi_layout"
    r = subprocess.run(["gcc", "-std=c11", "-I", str(ROOT / "include"), str(src), "-o", str(exe)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    got = [int(v) for v in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()]
    assert len(got) == len(expect)
    bad = [(what, py, c) for (what, py), c in zip(expect, got) if py != c]
    assert not bad, f"ctypes mirror != header (what, ctypes, C): {bad[:8]}"


def test_every_header_struct_has_a_mirror():
    import re
    text = (ROOT / "include" / "hoigen_b200.h").read_text()
    structs = set(re.findall(r"^\}\s*(hoigen_[a-z0-9_]+);", text, flags=re.M)) - {"hoigen_conv_op_kind"}
    assert structs == set(MIRRORS), structs ^ set(MIRRORS)
