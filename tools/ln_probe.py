"""Time the encoder's residual + LayerNorm pass (hoigen_add_layernorm768) at the bench shape; HOIGEN_LN_LEGACY=1 = old kernel."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hoigen_b200 import _cabi  # noqa: E402

dev = torch.device("cuda:0")
_cabi.init(dev)
M = 64 * 197
x = torch.randn(M, 768, device=dev)
d1 = torch.randn(M, 768, device=dev).bfloat16()
d2 = torch.randn(M, 768, device=dev).bfloat16()
g, b, cb = torch.ones(768, device=dev), torch.zeros(768, device=dev), torch.zeros(768, device=dev)
h = torch.empty(M, 768, device=dev, dtype=torch.bfloat16)
xb = torch.empty(M, 768, device=dev, dtype=torch.bfloat16)
flush = torch.empty(300 << 20, dtype=torch.uint8, device=dev)
for name, args in (("ln1 (x+=d1+d2+bias -> h)", (x, d1, d2, cb, g, b, h, None)), ("ln2 (x+=d1 -> h, xb)", (x, d1, None, None, g, b, h, xb))):
    ts = []
    for rep in range(14):
        flush.zero_()
        _cabi.profile(True)
        _cabi.call("hoigen_add_layernorm768", *[a.data_ptr() if a is not None else None for a in args], M)
        recs = _cabi.profile_read()
        _cabi.profile(False)
        if rep >= 4:
            ts.append(recs[-1][1] * 1e3)
    ts.sort()
    print(f"legacy={os.environ.get('HOIGEN_LN_LEGACY', '0')} {name}: median {ts[len(ts) // 2]:.1f} us, min {ts[0]:.1f}")
