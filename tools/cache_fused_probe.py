"""Time hoigen_score_cache_fused at the bench shape (Ktot = 7680, N = 4096, C = 117) — run once per HOIGEN_CF_DEBUG /
HOIGEN_CF_NSPLIT setting (the library reads them at launch)."""
import ctypes as C_
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hoigen_b200 import _cabi  # noqa: E402

dev = torch.device("cuda:0")
_cabi.init(dev)
ktot, N, C = int(os.environ.get("KTOT", 7680)), int(os.environ.get("NROWS", 4096)), int(os.environ.get("NCLS", 117))
g = torch.Generator().manual_seed(0)
B = 64
pair_off = torch.arange(0, ktot + 1, ktot // B, dtype=torch.int32, device=dev)[: B + 1].contiguous()
f = torch.randn(3, ktot, 512, generator=g).bfloat16().to(dev)
sw = _cabi.ScoreWeights()
sw.num_classes, sw.cache_rows, sw.affinity, sw.beta = C, N, 0, 5.0
keep = []
for x in range(3):
    t = [torch.randn(N, 512, generator=g).bfloat16().to(dev), (torch.rand(C, N, generator=g) < 0.02).bfloat16().to(dev),
         torch.zeros(N, device=dev), torch.zeros(C, device=dev), torch.ones(C, device=dev)]
    keep.append(t)
    sw.cache_keys[x], sw.label_t[x], sw.cache_bias[x] = t[0].data_ptr(), t[1].data_ptr(), t[2].data_ptr()
    sw.bias_term[x], sw.colscale[x] = t[3].data_ptr(), t[4].data_ptr()
img = torch.zeros(B, C, device=dev)
parts = torch.empty(int(_cabi.load().hoigen_cache_fused_workspace_bytes(ktot, C)) // 4, device=dev)
ld = (C + 3) // 4 * 4
logits = torch.empty(ktot, ld, device=dev)
bias_ptrs = (C_.c_void_p * 3)(*[keep[x][2].data_ptr() for x in range(3)])


def run():
    _cabi.call("hoigen_score_cache_fused", C_.byref(sw), f.data_ptr(), bias_ptrs, img.data_ptr(), pair_off.data_ptr(), B, ktot, 0, 5.0,
               parts.data_ptr(), logits.data_ptr(), ld)


for _ in range(3):
    run()
torch.cuda.synchronize()
_cabi.profile(True)
for _ in range(10):
    run()
recs = _cabi.profile_read()
_cabi.profile(False)
ms = sorted(r[1] for r in recs if r[0] == "cache_fused")
print(f"debug={os.environ.get('HOIGEN_CF_DEBUG', '0')} nsplit={os.environ.get('HOIGEN_CF_NSPLIT', 'auto')}: cache_fused median {ms[len(ms) // 2] * 1e3:.1f} us "
      f"(min {ms[0] * 1e3:.1f}); combine {sorted(r[1] for r in recs if r[0] == 'cache_combine')[5] * 1e3:.1f} us")
