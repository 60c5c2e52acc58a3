"""f1 measurement: batched association on the GPU (one launch per batch) vs the oracle port of the reference's per-image /
per-class Python loop on the host cores.  Detections/s over a batch of 64 images x ~1000 triplets."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pathlib import Path
import torch
from hoigen_b200 import synthetic as S
from hoigen_b200.evaluate import HOIAssociator
from oracle import eval_ref as E

dev = torch.device("cuda:0")
onv = json.load(open(Path(S.__file__).parent / "data" / "object_tables.json"))["hico_object_n_verb_to_interaction"]
conv = E.conversion_table(onv)
B = 64
dets = E.synthetic_detections(B, 5, 8, 8)
tgts = E.make_targets(dets, conv, seed=6, per_image=10)
ndet = sum(int(d["scores"].numel()) for d in dets)
assoc = HOIAssociator(onv)
ddev = [{k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in d.items()} for d in dets]
from hoigen_b200.evaluate import _pack
class _L(list): pass
dl = _L(ddev); dl.packed = _pack(ddev)
for _ in range(5): res = assoc(dl, tgts)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(50): res = assoc(dl, tgts)
torch.cuda.synchronize()
gpu_ms = (time.perf_counter() - t0) / 50 * 1e3
from hoigen_b200 import _cabi
_cabi.profile(True); assoc(dl, tgts); recs = _cabi.profile_read(); _cabi.profile(False)
k_us = [r[1] * 1e3 for r in recs if r[0] == "associate_pairs"]
t0 = time.perf_counter()
ref = E.associate_batch(dets, tgts, conv)
cpu_ms = (time.perf_counter() - t0) * 1e3
ok = all(torch.equal(a[2], b[2].cpu()) for a, b in zip(ref, res))
print(json.dumps({"images": B, "detections": ndet, "ground_truth_pairs": sum(int(t["hoi"].numel()) for t in tgts),
                  "gpu_call_ms": gpu_ms, "kernel_us": k_us, "cpu_port_ms": cpu_ms, "labels_equal": ok,
                  "detections_per_s_gpu": ndet / gpu_ms * 1e3, "detections_per_s_cpu": ndet / cpu_ms * 1e3}))
