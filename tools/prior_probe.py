"""Time hoigen_prior_tokens alone (50 back-to-back launches between two events): B images x n boxes."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hoigen_b200 import _cabi, synthetic as S  # noqa: E402

dev = torch.device("cuda:0")
B, nh, no = int(os.environ.get("B", 64)), 8, 8
head = S.make_head_state(117, 64)
props = S.make_region_props(B, nh, no, seed=5)
n_list = [p["boxes"].shape[0] for p in props]
n_max = max(n_list)
box_off = torch.tensor(np.concatenate([[0], np.cumsum(n_list)]), dtype=torch.int32, device=dev)
boxes = torch.cat([p["boxes"] for p in props]).to(dev).contiguous()
scores = torch.cat([p["scores"] for p in props]).to(dev).contiguous()
labels = torch.cat([p["labels"] for p in props]).to(dev).contiguous()
T = head.tensors
w = [T[f"priors_downproj.layers.{i}.weight"].t().contiguous().to(dev) for i in range(3)]
bb = [T[f"priors_downproj.layers.{i}.bias"].contiguous().to(dev) for i in range(3)]
oe = head.attrs["object_embedding"].to(dev).contiguous()
prior = torch.empty(B, n_max, 64, device=dev)
mask = torch.empty(B, n_max, device=dev, dtype=torch.uint8)
_cabi.init(dev)


def run():
    _cabi.call("hoigen_prior_tokens", boxes.data_ptr(), scores.data_ptr(), labels.data_ptr(), box_off.data_ptr(), oe.data_ptr(),
               w[0].data_ptr(), bb[0].data_ptr(), w[1].data_ptr(), bb[1].data_ptr(), w[2].data_ptr(), bb[2].data_ptr(),
               224.0, 224.0, B, n_max, oe.shape[0], prior.data_ptr(), mask.data_ptr())


for _ in range(5):
    run()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(50):
    run()
e1.record()
torch.cuda.synchronize()
print(f"prior_tokens B={B} n={n_max}: {e0.elapsed_time(e1) / 50 * 1e3:.1f} us per launch (back to back)")
