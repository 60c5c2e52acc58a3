"""Host-side pipelining probe: sequential forward vs launch-ahead, ms/step (device events + wall clock)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from hoigen_b200 import synthetic as S
from hoigen_b200.detector import UPT

dev = torch.device("cuda:0")
B = 64
enc, head = S.make_encoder_state(), S.make_head_state(117, 4096)
m = UPT.from_state(enc, head).to(dev).eval()
m.pack_weights(); m.clip_head.image_encoder.pack_weights()
imgs = [S.make_images(B, seed=i + 1).to(dev) for i in range(4)]
props = [[dict({k: v.to(dev) for k, v in S.make_boxes(64 * r + b, 8, 8).items()}, n_human=8) for b in range(B)] for r in range(4)]
dino = [S.make_dino_features(B, seed=7 + r).to(dev) for r in range(4)]

def seq(n):
    for i in range(n):
        m.forward_from_proposals(imgs[i % 4], props[i % 4], dino[i % 4])

def pipe(n):
    pend = m.launch_from_proposals(imgs[0], props[0], dino[0])
    for i in range(n):
        nxt = m.launch_from_proposals(imgs[(i + 1) % 4], props[(i + 1) % 4], dino[(i + 1) % 4]) if i + 1 < n else None
        m.finish(pend)
        pend = nxt

for name, fn in (("sequential", seq), ("launch-ahead", pipe), ("sequential", seq), ("launch-ahead", pipe)):
    fn(5)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    fn(20)
    torch.cuda.synchronize()
    print(f"{name}: {(time.perf_counter() - t0) * 1e3 / 20:.3f} ms/step", flush=True)

# host cost of a launch alone (GPU kept busy): time 10 launches without finishing
torch.cuda.synchronize()
t0 = time.perf_counter()
ps = [m.launch_from_proposals(imgs[i % 4], props[i % 4], dino[i % 4]) for i in range(10)]
t1 = time.perf_counter()
for p in ps:
    m.finish(p)
torch.cuda.synchronize()
print(f"host time per launch: {(t1 - t0) * 1e2:.3f} ms; finish of 10: {(time.perf_counter() - t1) * 1e3:.2f} ms")
