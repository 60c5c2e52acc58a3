"""Optional fast form of the reference's DINO branch (U:1616-1618: `dino_model(images_clip)` -> L2
normalise), for callers who run the WHOLE eval step and not only the four kernel groups of the path.

The branch is a stock torchvision ResNet-50 (fc = Identity) that the reference runs in fp32, ~160 small launches per
batch; measured next to the accelerated path it is the larger half of the step (bench.py `full_with_dino_r50`: 10.1 ms per
64-image step with the stock module against 3.7 ms without the branch).  `FastDinoR50` keeps the module and its weights
but changes how it is executed — every BatchNorm folded into its convolution (exact algebra in fp32, eval mode), bf16
channels-last tensors, the whole network replayed from ONE CUDA graph per batch size — i.e. cuDNN stays the engine: this
is library tuning outside the hand-written kernels, NOT part of the parity-gated path, and OFF unless asked for
(`UPT.accelerate_dino()`).  Outputs are the L2-normalised fp32 features the scoring stage takes (`dino_image_features`);
against the stock fp32 module they differ by bf16 rounding (tests/test_gpu_e2e.py::test_fast_dino_*).
"""
from __future__ import annotations

import copy
from typing import Dict

import torch
from torch import nn


def _fold_bn(conv: nn.Conv2d, bn: nn.BatchNorm2d) -> nn.Conv2d:
    """conv -> bn (eval) == conv' with w' = w * gamma / sqrt(var + eps), b' = beta + (b - mean) * gamma / sqrt(var + eps)."""
    w = conv.weight.detach().float()
    b = conv.bias.detach().float() if conv.bias is not None else torch.zeros(w.shape[0], device=w.device)
    scale = bn.weight.detach().float() / torch.sqrt(bn.running_var.detach().float() + bn.eps)
    fused = nn.Conv2d(conv.in_channels, conv.out_channels, conv.kernel_size, conv.stride, conv.padding, conv.dilation,
                      conv.groups, bias=True).to(w.device)
    fused.weight.data = w * scale[:, None, None, None]
    fused.bias.data = bn.bias.detach().float() + (b - bn.running_mean.detach().float()) * scale
    return fused


def fold_batchnorms(model: nn.Module) -> nn.Module:
    """A copy of a torchvision ResNet with every (conv, bn) pair folded and the bn replaced by Identity."""
    m = copy.deepcopy(model).eval()
    m.conv1, m.bn1 = _fold_bn(m.conv1, m.bn1), nn.Identity()
    for layer in (m.layer1, m.layer2, m.layer3, m.layer4):
        for blk in layer:
            blk.conv1, blk.bn1 = _fold_bn(blk.conv1, blk.bn1), nn.Identity()
            blk.conv2, blk.bn2 = _fold_bn(blk.conv2, blk.bn2), nn.Identity()
            blk.conv3, blk.bn3 = _fold_bn(blk.conv3, blk.bn3), nn.Identity()
            if blk.downsample is not None:
                blk.downsample = nn.Sequential(_fold_bn(blk.downsample[0], blk.downsample[1]))
    return m


class FastDinoR50(nn.Module):
    """`features = fast(images)` -> (B, 2048) fp32, L2-normalised (U:1617-1618 in one call)."""

    def __init__(self, dino_model: nn.Module, use_graph: bool = True):
        super().__init__()
        dev = next(dino_model.parameters()).device
        if dev.type != "cuda":
            raise ValueError("FastDinoR50 needs the module on a CUDA device")
        self.net = fold_batchnorms(dino_model).to(device=dev, dtype=torch.bfloat16, memory_format=torch.channels_last)
        self.use_graph = use_graph
        self._graphs: Dict[tuple, tuple] = {}

    @torch.no_grad()
    def _run(self, x: torch.Tensor) -> torch.Tensor:
        f = self.net(x.to(dtype=torch.bfloat16, memory_format=torch.channels_last)).float()
        return f / f.norm(dim=-1, keepdim=True)

    @torch.no_grad()
    def forward(self, images: torch.Tensor) -> torch.Tensor:
        if not self.use_graph:
            return self._run(images)
        # one graph (and one pair of static buffers) per (batch size, CUDA stream): forwards on different streams may overlap
        B = (int(images.shape[0]), torch.cuda.current_stream(images.device).cuda_stream)
        entry = self._graphs.get(B)
        if entry is None:
            static_in = torch.empty_like(images, dtype=torch.float32)
            side = torch.cuda.Stream(device=images.device)
            side.wait_stream(torch.cuda.current_stream(images.device))
            with torch.cuda.stream(side):
                static_in.copy_(images)
                for _ in range(3):                      # cuDNN algorithm selection happens outside the capture
                    self._run(static_in)
            torch.cuda.current_stream(images.device).wait_stream(side)
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                static_out = self._run(static_in)
            entry = self._graphs[B] = (graph, static_in, static_out)
        graph, static_in, static_out = entry
        static_in.copy_(images)
        graph.replay()
        return static_out.clone()
