"""Time DETR's ResNet-50 backbone body (FrozenBatchNorm, layer4 only; f3's backbone half) at a DETR-sized padded batch: the stock
fp32 torchvision module against the repo's convolution kernels (hoigen_b200.dino.KernelDetrBackboneBody)."""
import os
import sys

import torch
import torchvision
from torchvision.models._utils import IntermediateLayerGetter
from torchvision.ops.misc import FrozenBatchNorm2d

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hoigen_b200.dino import KernelDetrBackboneBody  # noqa: E402

dev = torch.device("cuda:0")
B, H, W = int(os.environ.get("B", 8)), int(os.environ.get("H", 800)), int(os.environ.get("W", 1216))
torch.manual_seed(0)
r50 = torchvision.models.resnet50(weights=None, norm_layer=FrozenBatchNorm2d)
body = IntermediateLayerGetter(r50, return_layers={"layer4": "0"}).to(dev).eval()
x = torch.randn(B, 3, H, W, device=dev)
fast = KernelDetrBackboneBody(body)


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
    t_stock = timeit(lambda: body(x))
    t_fast = timeit(lambda: fast(x))
    ref, got = body(x)["0"], fast(x)["0"]
cos = torch.nn.functional.cosine_similarity(got, ref, dim=1).min().item()
gflop = 4.09 * H * W / (224 * 224) * 2          # ~ 4.09 GMAC at 224 x 224
print(f"B={B} {H}x{W}: stock fp32 {t_stock:.2f} ms, own kernels {t_fast:.2f} ms ({gflop * B / t_fast:.0f} TFLOP/s), "
      f"layer4 {tuple(ref.shape)}, min cosine {cos:.6f}")
