"""Time the DETR detector of the proposal stage (U:1594-1599) at a DETR-sized padded batch: stock fp32 modules (torchvision
FrozenBatchNorm ResNet-50 body + nn.MultiheadAttention transformer, evaluated the way the reference evaluates them) against
hoigen_b200.detr.KernelDetr; then the per-launch profile of the kernel path."""
import os
import sys
from collections import defaultdict

import torch
import torchvision
from torchvision.models._utils import IntermediateLayerGetter
from torchvision.ops.misc import FrozenBatchNorm2d

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hoigen_b200 import _cabi  # noqa: E402
from hoigen_b200.detr import KernelDetr  # noqa: E402
from oracle import detr_ref as D  # noqa: E402  (stands in for the reference's detector object: same modules, same names)

dev = torch.device("cuda:0")
B, H, W = int(os.environ.get("B", 8)), int(os.environ.get("H", 800)), int(os.environ.get("W", 1216))
torch.manual_seed(0)
det = D.DetrRef().eval()
D.seeded_state(det, 17)
r50 = torchvision.models.resnet50(weights=None, norm_layer=FrozenBatchNorm2d)


class BackboneBase(torch.nn.Module):
    def __init__(self):
        super().__init__()
        self.body = IntermediateLayerGetter(r50, return_layers={"layer4": "0"})


det.backbone = torch.nn.Sequential(BackboneBase(), torch.nn.Identity())
det = det.to(dev).eval()
images = torch.randn(B, 3, H, W, device=dev)
mask = torch.zeros(B, H, W, dtype=torch.bool, device=dev)
mask[1::2, :, W - 160:] = True
fast = KernelDetr(det)


def stock():
    feat = det.backbone[0].body(images)["0"]
    m = torch.nn.functional.interpolate(mask[None].float(), size=feat.shape[-2:]).to(torch.bool)[0]
    return det.forward_features(feat, m)


def timeit(fn, n=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


with torch.no_grad():
    t_stock = timeit(stock)
    t_fast = timeit(lambda: fast(images, mask))
    rl, rb = stock()
    fl, fb = fast(images, mask)
print(f"B={B} {H}x{W}: stock fp32 {t_stock:.2f} ms, own kernels {t_fast:.2f} ms; logits max-abs {(fl - rl).abs().max().item():.3e} "
      f"(|ref| <= {rl.abs().max().item():.2f}), boxes {(fb - rb).abs().max().item():.3e}")
_cabi.profile(True)
for _ in range(3):
    fast(images, mask)
recs = _cabi.profile_read()
_cabi.profile(False)
agg = defaultdict(list)
for r in recs:
    agg[r[0]].append(r[1])
tot = sum(sum(v) for v in agg.values()) / 3
print(f"profiled (serialised) sum per forward: {tot:.3f} ms")
for tag, v in sorted(agg.items(), key=lambda kv: -sum(kv[1]))[:14]:
    print(f"  {tag:28s} x{len(v) // 3:3d}  {sum(v) / 3 * 1e3:8.1f} us  ({sum(v) / 3 / tot * 100:4.1f} %)")
