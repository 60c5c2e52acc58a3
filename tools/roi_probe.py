"""Time the RoI stage (hoigen_roi_pair_features: axis weights, RoIAlign+mean, pair assembly) at the bench shape; HOIGEN_ROI_SIMT=1 =
the fp32 SIMT feature kernel instead of the tensor-core one."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hoigen_b200 import _cabi, synthetic as S  # noqa: E402

dev = torch.device("cuda:0")
_cabi.init(dev)
B, nh, no = int(os.environ.get("B", 64)), int(os.environ.get("NH", 8)), int(os.environ.get("NO", 8))
props = S.make_region_props(B, nh, no)
n_list = [p["boxes"].shape[0] for p in props]
k_list = [nh * (n - 1) for n in n_list]
box_off = torch.tensor(np.concatenate([[0], np.cumsum(n_list)]), dtype=torch.int32, device=dev)
pair_off = torch.tensor(np.concatenate([[0], np.cumsum(k_list)]), dtype=torch.int32, device=dev)
ntot, ktot = sum(n_list), sum(k_list)
tokens = torch.randn(B * 197, 512, device=dev)
boxes = torch.cat([p["boxes"] for p in props]).to(dev).contiguous()
single, union = torch.empty(ntot, 512, device=dev), torch.empty(ktot, 512, device=dev)
pb = torch.empty(3, ktot, 512, device=dev, dtype=torch.bfloat16)
ws = torch.empty((ntot + ktot) * 32, device=dev)


def run():
    _cabi.call("hoigen_roi_pair_features", tokens.data_ptr(), boxes.data_ptr(), box_off.data_ptr(), pair_off.data_ptr(), B, ntot, ktot,
               14.0 / 224.0, ws.data_ptr(), single.data_ptr(), union.data_ptr(), pb.data_ptr(), None)


for _ in range(3):
    run()
torch.cuda.synchronize()
_cabi.profile(True)
for _ in range(10):
    run()
recs = _cabi.profile_read()
_cabi.profile(False)
out = []
for tag in ("roi_weights", "roi_features", "pair_assemble"):
    if not any(r[0] == tag for r in recs):
        continue
    ms = sorted(r[1] for r in recs if r[0] == tag)
    out.append(f"{tag} {ms[len(ms) // 2] * 1e3:.1f} us")
print(f"simt={os.environ.get('HOIGEN_ROI_SIMT', '0')} B={B} boxes={nh}+{no}: " + ", ".join(out))
