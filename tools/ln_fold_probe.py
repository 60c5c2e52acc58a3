"""A/B of the LayerNorm-folded GEMM epilogue against the plain one on the encoder's QKV / c_fc shapes (same process,
interleaved launches, library per-launch CUDA events)."""
import math
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hoigen_b200 import _cabi  # noqa: E402

dev = torch.device("cuda:0")
_cabi.init(dev)
M, K = 12608, 768
g = torch.Generator().manual_seed(0)
a = torch.randn(M, K, generator=g).bfloat16().to(dev)
stats = torch.rand(M, 2, generator=g).to(dev)
for N, act in ((2304, 0), (3072, 1)):
    w = (torch.randn(N, K, generator=g) / math.sqrt(K)).bfloat16().to(dev)
    bias = torch.randn(N, generator=g).to(dev)
    colsum = torch.randn(N, generator=g).to(dev)
    out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    res = {"plain": [], "ln": []}
    for rep in range(24):
        for kind in ("plain", "ln"):
            flush.zero_()
            _cabi.profile(True)
            if kind == "plain":
                _cabi.gemm_bf16(a, w, bias=bias, act=act, out_bf16=out)
            else:
                _cabi.gemm_bf16(a, w, bias=bias, act=act, out_bf16=out, ln_stats=stats, ln_colsum=colsum)
            recs = _cabi.profile_read()
            _cabi.profile(False)
            if rep >= 4:
                res[kind].append(recs[-1][1] * 1e3)
    for kind, v in res.items():
        v.sort()
        print(f"N={N} act={act} {kind:5s}: median {v[len(v) // 2]:.1f} us  min {v[0]:.1f}  ({2.0 * M * N * K / v[len(v) // 2] / 1e6:.0f} TFLOP/s)")
