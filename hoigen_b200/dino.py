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

`KernelDinoR50` (round 2) is the same network on the repo's OWN kernels: every convolution runs on the tcgen05 GEMM of
`csrc/gemm.cu` -- 1x1 convolutions are plain GEMMs over NHWC rows, 3x3 / stride-1 convolutions are nine accumulated
products of row-shifted views of the same activation matrix (an implicit GEMM: the activations carry a one-pixel zero halo,
so no im2col matrix exists), bias + identity + ReLU + halo zeroing live in the GEMM epilogue -- plus four small row kernels
(`csrc/conv_rows.cu`: stem im2col, max-pool, the stride-2 gathers, average-pool + L2 norm).  One C call
(`hoigen_conv_plan_run`) launches the whole 60-step plan.
"""
from __future__ import annotations

import copy
import ctypes as C
from typing import Dict, List

import torch
from torch import nn

from . import _cabi


def _fold_bn(conv: nn.Conv2d, bn: nn.Module) -> nn.Conv2d:
    """conv -> bn (eval) == conv' with w' = w * gamma / sqrt(var + eps), b' = beta + (b - mean) * gamma / sqrt(var + eps)."""
    w = conv.weight.detach().float()
    b = conv.bias.detach().float() if conv.bias is not None else torch.zeros(w.shape[0], device=w.device)
    scale = bn.weight.detach().float() / torch.sqrt(bn.running_var.detach().float() + getattr(bn, "eps", 1e-5))   # DETR's FrozenBatchNorm2d: 1e-5
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


class KernelResNet50(nn.Module):
    """A torchvision ResNet-50 (eval; BatchNorm2d or DETR's FrozenBatchNorm2d) executed on the repo's kernels for any image
    size: `rows, (h, w) = net.features(images)` gives layer4 as haloed NHWC bf16 rows of (B, h + 2, w + 2, 2048).

    Weights: the BatchNorm-folded convolutions, bf16, packed K-major: 1x1 -> (Cout, Cin); 3x3 -> (Cout, 9 Cin) with
    k = (ky * 3 + kx) * Cin + c; stem 7x7 -> (64, 160), k = (ky * 7 + kx) * 3 + c (im2col form) and (64, 192),
    k = ky * 24 + kx * 3 + c (fused 224 x 224 stem kernel).
    """

    STEM_K = 160

    def __init__(self, resnet: nn.Module):
        super().__init__()
        dev = next(iter(resnet.parameters())).device
        if dev.type != "cuda":
            raise ValueError("KernelResNet50 needs the module on a CUDA device")
        m = fold_batchnorms(resnet)
        self.device_ = dev
        self._keep: List[torch.Tensor] = []

        def pack(conv: nn.Conv2d):
            w = conv.weight.detach().float()
            co, ci, kh, kw = w.shape
            wk = w.permute(0, 2, 3, 1).reshape(co, kh * kw * ci)
            if kh == 7:
                wk = torch.nn.functional.pad(wk, (0, self.STEM_K - wk.shape[1]))
            wk = wk.to(torch.bfloat16).contiguous()
            b = conv.bias.detach().float().contiguous()
            self._keep += [wk, b]
            return wk, b

        self.stem = pack(m.conv1)                      # (64, 160): the two-step form (im2col + GEMM), any image size
        w7 = m.conv1.weight.detach().float()           # (64, 3, 7, 7) -> (64, 192): column = ky * 24 + kx * 3 + c
        w_runs = torch.zeros(64, 7, 24, device=dev)
        w_runs[:, :, :21] = w7.permute(0, 2, 3, 1).reshape(64, 7, 21)
        self.stem_fused = (torch.nn.functional.pad(w_runs.reshape(64, 168), (0, 24)).to(torch.bfloat16).contiguous(), self.stem[1])
        self._keep.append(self.stem_fused[0])
        self.fused_stem = True                         # False: the two-step form (im2col + GEMM)
        self.merge_downsample = True                   # False: downsample as its own GEMM + bf16 identity epilogue
        self.implicit_stride2 = True                   # False: stride-2 3x3 convolutions through the nine-tap gather (im2col)
        self.stages = []
        for layer in (m.layer1, m.layer2, m.layer3, m.layer4):
            blocks = []
            for blk in layer:
                if blk.conv2.dilation[0] != 1 or blk.conv1.stride[0] != 1:
                    raise NotImplementedError("KernelResNet50: dilated / stride-on-conv1 bottlenecks are not supported")
                d = dict(c1=pack(blk.conv1), c2=pack(blk.conv2), c3=pack(blk.conv3),
                         ds=pack(blk.downsample[0]) if blk.downsample is not None else None,
                         stride=blk.conv2.stride[0], c3ds=None)
                if d["ds"] is not None:
                    # conv3(t) + downsample(x) as ONE product: [t | x] . [W3 | Wds]^T + (b3 + bds)  (hoigen_gemm_params.a2)
                    d["c3ds"] = (torch.cat([d["c3"][0], d["ds"][0]], dim=1).contiguous(), (d["c3"][1] + d["ds"][1]).contiguous())
                    self._keep += list(d["c3ds"])
                blocks.append(d)
            self.stages.append(blocks)
        self._plans: Dict[tuple, dict] = {}

    # ---- plan construction: buffers + the op list for one (batch, image size, head, stream) ----
    def _build_plan(self, B: int, Hi: int, Wi: int, head: str) -> dict:
        dev = self.device_
        bf = dict(device=dev, dtype=torch.bfloat16)
        ops: List[_cabi.ConvOp] = []
        bufs: List[torch.Tensor] = []

        pool: Dict[tuple, list] = {}

        def new(rows, cols):
            # every buffer is written in full by its producer (halo rows included) and the plan runs in order on one stream,
            # so a buffer whose last reader has been enqueued can back a later tensor of the same shape
            free_list = pool.get((rows, cols))
            if free_list:
                return free_list.pop()
            t = torch.zeros(rows, cols, **bf)
            bufs.append(t)
            return t

        def release(*ts):
            for t in ts:
                if t is not None:
                    pool.setdefault((t.shape[0], t.shape[1]), []).append(t)

        def gemm(a, wb, out, *, relu, halo=None, taps=0, res=None, a2=None, stride=1):
            op = _cabi.ConvOp()
            op.kind = _cabi.CONV_OP_GEMM
            n_out = wb[0].shape[0]
            # CTA-pair tiles with the TMA-store epilogue whenever the output is wide enough (the tile chooser's cost model
            # is tuned for K >= 768; these products are short-K and store-bound)
            bn = 2256 if n_out >= 256 else (2128 if (n_out >= 128 or a.shape[1] != self.STEM_K) else 0)
            op.gemm = _cabi.gemm_params(a, wb[0], bias=wb[1], act=_cabi.ACT_RELU if relu else _cabi.ACT_NONE, out_bf16=out,
                                        conv_taps=taps, halo=halo, res_bf16=res, block_n=bn, a2=a2, conv_stride=stride)
            ops.append(op)

        def rowop(kind, src, dst, h=0, w=0, c=0, taps=0):
            op = _cabi.ConvOp()
            op.kind, op.in_, op.out = kind, src.data_ptr() if src is not None else 0, dst.data_ptr()
            op.batch, op.h, op.w, op.c, op.taps = B, h, w, c, taps
            ops.append(op)

        half = lambda v: (v + 1) // 2
        H1, W1 = half(Hi), half(Wi)                    # stem output
        stem_out = new(B * H1 * W1, 64)
        if self.fused_stem:
            op = _cabi.ConvOp()
            op.kind, op.out, op.batch, op.h, op.w = _cabi.CONV_OP_STEM_CONV, stem_out.data_ptr(), B, Hi, Wi
            op.gemm.w, op.gemm.bias = self.stem_fused[0].data_ptr(), self.stem_fused[1].data_ptr()
            ops.append(op)
        else:
            stem_rows = new(B * H1 * W1, self.STEM_K)
            rowop(_cabi.CONV_OP_STEM_IM2COL, None, stem_rows, Hi, Wi)
            gemm(stem_rows, self.stem, stem_out, relu=True)
            release(stem_rows)
        H, W = half(H1), half(W1)                      # after the max-pool
        x = new(B * (H + 2) * (W + 2), 64)
        rowop(_cabi.CONV_OP_MAXPOOL, stem_out, x, H1, W1, 64)
        release(stem_out)
        cin = 64
        for blocks in self.stages:
            for blk in blocks:
                width = blk["c1"][0].shape[0]
                cout = blk["c3"][0].shape[0]
                halo_in = (H + 2, W + 2)
                rows_in = B * halo_in[0] * halo_in[1]
                t1 = new(rows_in, width)
                gemm(x, blk["c1"], t1, relu=True, halo=halo_in)
                if blk["stride"] == 2:
                    Ho, Wo = half(H), half(W)
                    halo_out = (Ho + 2, Wo + 2)
                    rows_out = B * halo_out[0] * halo_out[1]
                    t2 = new(rows_out, width)
                    if self.implicit_stride2 and width % 64 == 0:
                        # four-phase split of t1 in the output's haloed geometry: every tap is a row shift of one phase block
                        g2 = new(4 * rows_out, width)
                        rowop(_cabi.CONV_OP_GATHER_S2, t1, g2, H, W, width, 4)
                        gemm(g2, blk["c2"], t2, relu=True, halo=halo_out, taps=9, stride=2)
                    else:
                        g2 = new(rows_out, 9 * width)
                        rowop(_cabi.CONV_OP_GATHER_S2, t1, g2, H, W, width, 9)
                        gemm(g2, blk["c2"], t2, relu=True, halo=halo_out)
                    release(g2)
                    merged = self.merge_downsample and rows_out > 128
                    gs = new(rows_out, cin)
                    rowop(_cabi.CONV_OP_GATHER_S2, x, gs, H, W, cin, 1)
                    if merged:
                        sc, x2 = None, gs
                    else:
                        sc = new(rows_out, cout)
                        gemm(gs, blk["ds"], sc, relu=False, halo=halo_out)
                        release(gs)
                    H, W = Ho, Wo
                else:
                    halo_out, rows_out = halo_in, rows_in
                    merged = self.merge_downsample and blk["ds"] is not None and rows_out > 128
                    t2 = new(rows_out, width)
                    gemm(t1, blk["c2"], t2, relu=True, halo=halo_out, taps=9)
                    if merged:
                        sc, x2 = None, x
                    elif blk["ds"] is not None:
                        sc = new(rows_out, cout)
                        gemm(x, blk["ds"], sc, relu=False, halo=halo_out)
                    else:
                        sc = x
                y = new(rows_out, cout)
                if merged:
                    gemm(t2, blk["c3ds"], y, relu=True, halo=halo_out, a2=x2)
                    release(t1, t2, x, x2 if x2 is not x else None)
                else:
                    gemm(t2, blk["c3"], y, relu=True, halo=halo_out, res=sc)
                    release(t1, t2, x, sc if sc is not x else None)
                x, cin = y, cout
        out = None
        if head == "dino":
            out = torch.empty(B, cin, device=dev, dtype=torch.float32)
            op = _cabi.ConvOp()
            op.kind, op.in_, op.out = _cabi.CONV_OP_AVGPOOL_L2NORM, x.data_ptr(), out.data_ptr()
            op.batch, op.h, op.w, op.c = B, H, W, cin
            ops.append(op)
        arr = (_cabi.ConvOp * len(ops))(*ops)
        return dict(ops=arr, n=len(ops), bufs=bufs, out=out, rows=x, hw=(H, W), channels=cin)

    def _run(self, images: torch.Tensor, head: str) -> dict:
        if images.device != self.device_ or images.dtype != torch.float32 or images.dim() != 4 or images.shape[1] != 3:
            raise ValueError("KernelResNet50 takes (B, 3, H, W) fp32 images on the module's device")
        images = images.contiguous()
        lib = _cabi.init(images.device)
        key = (int(images.shape[0]), int(images.shape[2]), int(images.shape[3]), head,
               torch.cuda.current_stream(images.device).cuda_stream)
        plan = self._plans.get(key)
        if plan is None:
            if len(self._plans) >= 8:                  # DETR batches change size: keep the activation sets bounded
                self._plans.clear()
            plan = self._plans[key] = self._build_plan(key[0], key[1], key[2], head)
        plan["ops"][0].in_ = images.data_ptr()
        _cabi.check(lib.hoigen_conv_plan_run(plan["ops"], plan["n"], _cabi.stream_ptr()), "hoigen_conv_plan_run")
        return plan

    @torch.no_grad()
    def features(self, images: torch.Tensor):
        """-> (rows, (h, w)): layer4 as haloed NHWC bf16 rows of (B, h + 2, w + 2, 2048), a workspace view valid until the
        next call with the same batch / image size on this stream."""
        plan = self._run(images, "features")
        return plan["rows"], plan["hw"]

    @torch.no_grad()
    def features_nchw(self, images: torch.Tensor) -> torch.Tensor:
        """-> layer4 (B, 2048, h, w) fp32, what `IntermediateLayerGetter(resnet, {'layer4': '0'})` returns."""
        rows, (h, w) = self.features(images)
        B = images.shape[0]
        return rows.view(B, h + 2, w + 2, -1)[:, 1:-1, 1:-1].permute(0, 3, 1, 2).float().contiguous()


class KernelDinoR50(KernelResNet50):
    """`features = net(images)` -> (B, 2048) fp32, L2-normalised (U:1617-1618), on the hand-written kernels."""

    def __init__(self, dino_model: nn.Module):
        super().__init__(dino_model)

    @torch.no_grad()
    def forward(self, images: torch.Tensor) -> torch.Tensor:
        if tuple(images.shape[1:]) != (3, 224, 224):
            raise ValueError("KernelDinoR50 takes (B, 3, 224, 224) fp32 images on the module's device")
        return self._run(images, "dino")["out"].clone()


class KernelDetrBackboneBody(nn.Module):
    """Drop-in for `detector.backbone[0].body` (torchvision IntermediateLayerGetter over a FrozenBatchNorm ResNet-50 returning
    {'0': layer4}; detr/models/backbone.py:69,72-73, call site U:1594): the same feature map from the repo's kernels."""

    def __init__(self, body: nn.Module):
        super().__init__()
        self.net = KernelResNet50(body)

    @torch.no_grad()
    def forward(self, images: torch.Tensor):
        from collections import OrderedDict
        return OrderedDict([("0", self.net.features_nchw(images.float()))])
