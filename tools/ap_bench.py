"""f2 measurement: DetectionAPMeter.eval on the GPU (two library sorts + one kernel for all 600 classes) vs the oracle
port of the reference's per-class host loop, on a sweep of 4 M detections."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from hoigen_b200 import _cabi
from hoigen_b200.evaluate import DetectionAPMeter
from oracle import eval_ref as E

dev = torch.device("cuda:0")
stream, num_gt = E.synthetic_meter_stream(5, num_cls=600, batches=40, per_batch=100000)
n = sum(int(s[0].numel()) for s in stream)
dstream = [(a.to(dev), b.to(dev), c.to(dev)) for a, b, c in stream]

def run():
    m = DetectionAPMeter(600, num_gt=num_gt)
    for a, b, c in dstream:
        m.append(a, b, c)
    return m.eval(), m
for _ in range(2): run()
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(5): ap, m = run()
torch.cuda.synchronize(); gpu_ms = (time.perf_counter() - t0) / 5 * 1e3
_cabi.profile(True); run(); recs = _cabi.profile_read(); _cabi.profile(False)
k_us = [r[1] * 1e3 for r in recs if r[0] == "ap_11point"]
t0 = time.perf_counter()
sc, lb = E.group_by_class(stream, 600)
ref_ap, _ = E.ap_11point(sc, lb, num_gt)
cpu_ms = (time.perf_counter() - t0) * 1e3
print(json.dumps({"detections": n, "classes": 600, "gpu_append_plus_eval_ms": gpu_ms, "kernel_us": k_us, "cpu_port_ms": cpu_ms,
                  "ap_equal": bool(torch.equal(ap.cpu(), ref_ap)), "mAP": float(ap.mean())}))
