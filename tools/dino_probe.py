"""Time the ResNet-50 branch (row a8) three ways on one GPU: the stock fp32 torchvision module, the cuDNN bf16 channels-last
graph (FastDinoR50) and the repo's own kernels (KernelDinoR50); then the per-launch profile of the kernel plan."""
import os
import sys
from collections import defaultdict

import torch
import torchvision

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hoigen_b200 import _cabi  # noqa: E402
from hoigen_b200.dino import FastDinoR50, KernelDinoR50  # noqa: E402

dev = torch.device("cuda:0")
B = int(os.environ.get("B", 64))
torch.manual_seed(0)
r50 = torchvision.models.resnet50(weights=None)
r50.fc = torch.nn.Identity()
r50 = r50.to(dev).eval()
imgs = torch.randn(B, 3, 224, 224, device=dev)


def timeit(fn, n=10):
    for _ in range(3):
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
    t_stock = timeit(lambda: r50(imgs))
    fast = FastDinoR50(r50)
    t_cudnn = timeit(lambda: fast(imgs))
    kern = KernelDinoR50(r50)
    t_kern = timeit(lambda: kern(imgs))
    split = KernelDinoR50(r50)
    split.merge_downsample = False                     # A/B: downsample as its own GEMM + identity epilogue
    t_split = timeit(lambda: split(imgs))
    print(f"conv3 + downsample merged {t_kern:.3f} ms, separate {t_split:.3f} ms, max-abs feature difference "
          f"{(kern(imgs) - split(imgs)).abs().max().item():.3e}")
    del split
    im2col = KernelDinoR50(r50)
    im2col.implicit_stride2 = False                    # A/B: stride-2 3x3 convolutions through the nine-tap gather
    t_g9 = timeit(lambda: im2col(imgs))
    print(f"stride-2 3x3 as implicit GEMM over the four-phase split {t_kern:.3f} ms, nine-tap gather + GEMM {t_g9:.3f} ms, "
          f"max-abs feature difference {(kern(imgs) - im2col(imgs)).abs().max().item():.3e}")
    del im2col
    ref = r50(imgs)
    ref = ref / ref.norm(dim=-1, keepdim=True)
    cos = (kern(imgs) * ref).sum(-1).min().item()
flops = 8.2e9 * B
print(f"B={B}: stock fp32 {t_stock:.3f} ms, cuDNN bf16 graph {t_cudnn:.3f} ms, own kernels {t_kern:.3f} ms "
      f"({flops / t_kern / 1e9:.0f} TFLOP/s on 8.2 GFLOP/img), min cosine vs stock {cos:.6f}")
_cabi.profile(True)
for _ in range(5):
    kern(imgs)
recs = _cabi.profile_read()
_cabi.profile(False)
agg = defaultdict(list)
for r in recs:
    agg[r[0]].append(r[1])
tot = sum(sum(v) for v in agg.values()) / 5
print(f"profiled (serialised) sum per forward: {tot:.3f} ms")
for tag, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
    print(f"  {tag:28s} x{len(v) // 5:3d}  {sum(v) / 5 * 1e3:8.1f} us  ({sum(v) / 5 / tot * 100:4.1f} %)  avg {sum(v) / len(v) * 1e3:6.1f} us")
