"""Diagnostic: does a launch/finish serving loop leave CYCLIC garbage (objects only the Python GC can free) that holds
device memory?  Runs steps with the GC off, then collects with DEBUG_SAVEALL and lists what was unreachable."""
import collections, gc, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from hoigen_b200 import _cabi, synthetic as S
from hoigen_b200.detector import UPT

dev = torch.device("cuda:0")
_cabi.init(dev)
B = 64
m = UPT.from_state(S.make_encoder_state(0), S.make_head_state(117, 4096)).to(dev)
imgs = S.make_images(B, seed=1).to(dev)
props = [dict({k: v.to(dev) for k, v in S.make_boxes(b, 8, 8).items()}, n_human=8) for b in range(B)]
dino = S.make_dino_features(B).to(dev)
streams = [torch.cuda.Stream(), torch.cuda.Stream()]

def step(i):
    with torch.cuda.stream(streams[i % 2]):
        return m.launch_from_proposals(imgs, props, dino)

for i in range(6):
    m.finish(step(i))
torch.cuda.synchronize()
gc.collect(); gc.disable()
seg0 = torch.cuda.memory_stats(dev)["segment.all.allocated"]; res0 = torch.cuda.memory_reserved(dev)
pend = step(0); keep = [None, None]
for i in range(40):
    nxt = step(i + 1)
    keep[i % 2] = m.finish(pend)
    pend = nxt
m.finish(pend); keep = None; pend = nxt = None
torch.cuda.synchronize()
seg1 = torch.cuda.memory_stats(dev)["segment.all.allocated"]; res1 = torch.cuda.memory_reserved(dev)
gc.set_debug(gc.DEBUG_SAVEALL)
n = gc.collect()
types = collections.Counter(type(o).__name__ for o in gc.garbage)
tens = sum(o.numel() * o.element_size() for o in gc.garbage if isinstance(o, torch.Tensor) and o.is_cuda)
print(f"segments {seg0} -> {seg1}, reserved {res0 >> 20} -> {res1 >> 20} MiB; unreachable objects {n}; cuda tensor bytes in garbage {tens >> 20} MiB")
print(types.most_common(12))
