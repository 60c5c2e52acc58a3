"""f3 measurement: the proposal stage (batched_nms + threshold + min/max instances) of a 64-image batch of 100 DETR
queries as ONE kernel (UPT.prepare_region_proposals_batched) vs the reference's per-image torch form on the same GPU
(UPT.prepare_region_proposals, torchvision's CUDA nms) and vs the per-image form on the host cores."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from hoigen_b200 import _cabi, synthetic as S
from hoigen_b200.detector import UPT
from oracle import hoi_forward_ref as O

dev = torch.device("cuda:0")
B, Q = 64, 100
m = UPT.from_state(S.make_encoder_state(0), S.make_head_state(117, 128, seed=2)).to(dev)
host = O.synthetic_detr_results(B, 7, Q)
res = [{k: v.to(dev) for k, v in r.items()} for r in host]

def timed(fn, n):
    for _ in range(3): fn()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): out = fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e3, out

batched_ms, got = timed(lambda: m.prepare_region_proposals_batched(res), 50)
torch_ms, ref = timed(lambda: m.prepare_region_proposals(res), 5)
t0 = time.perf_counter(); cpu = O.prepare_region_proposals(host, 0, 0.2, 3, 15); cpu_ms = (time.perf_counter() - t0) * 1e3
_cabi.profile(True); m.prepare_region_proposals_batched(res); recs = _cabi.profile_read(); _cabi.profile(False)
k_us = [r[1] * 1e3 for r in recs if r[0] == "prepare_proposals"]
boxes, scores, labels, n_list, nh_list = got
same = torch.equal(boxes, torch.cat([r["boxes"] for r in ref])) and torch.equal(labels, torch.cat([r["labels"] for r in ref])) \
    and torch.equal(boxes.cpu(), torch.cat([r["boxes"] for r in cpu]))
print(json.dumps({"images": B, "queries": Q, "kept_boxes": int(sum(n_list)), "batched_ms": batched_ms, "kernel_us": k_us,
                  "per_image_torch_gpu_ms": torch_ms, "per_image_torch_cpu_ms": cpu_ms, "cores": os.cpu_count(), "equal": bool(same)}))
