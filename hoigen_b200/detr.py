"""The DETR detector behind UPT.forward's proposal stage (U:1594-1599) on this repo's kernels — opt-in
(`UPT.accelerate_detr()`), bf16 operands, fp32 residual stream.

    features, pos = detector.backbone(nested)                       # ResNet-50 body -> dino.KernelResNet50 (conv plan)
    hs = detector.transformer(detector.input_proj(src), mask, detector.query_embed.weight, pos)
    logits, boxes = detector.class_embed(hs), detector.bbox_embed(hs).sigmoid()

Every nn.Linear / 1x1 convolution is `hoigen_gemm_bf16` (tcgen05); between them run two row kernels (`csrc/detr_rows.cu`):
`hoigen_add_layernorm256` (residual + post-norm, emitting the bf16 operands of the next products incl. `x + pos`) and
`hoigen_attention_heads32` (nn.MultiheadAttention's core for head_dim 32 with the key-padding mask, online softmax).  Only the
LAST decoder layer's normalised activations are produced — all UPT.forward reads (U:1604: `outputs_class[-1]`).

Token layout: the backbone's haloed NHWC rows are used as they are — the halo pixels simply become padding keys (mask = 1), so
no gather sits between the backbone and the transformer.  The sine position encoding (detr/models/position_encoding.py:28-50) is
a handful of torch ops on the (B, h, w) mask per batch.
"""
from __future__ import annotations

import math
from typing import Dict

import torch
from torch import nn
import torch.nn.functional as F

from . import _cabi
from .dino import KernelResNet50


def sine_position_embedding(mask: torch.Tensor, num_pos_feats: int = 128, temperature: float = 10000.0) -> torch.Tensor:
    """mask (B, h, w) bool (True = padding) -> (B, h, w, 2 * num_pos_feats) fp32: PositionEmbeddingSine(normalize=True), y half first."""
    not_mask = ~mask
    y = not_mask.cumsum(1, dtype=torch.float32)
    x = not_mask.cumsum(2, dtype=torch.float32)
    y = y / (y[:, -1:, :] + 1e-6) * (2 * math.pi)
    x = x / (x[:, :, -1:] + 1e-6) * (2 * math.pi)
    dim_t = torch.arange(num_pos_feats, dtype=torch.float32, device=mask.device)
    dim_t = temperature ** (2 * (dim_t // 2) / num_pos_feats)
    px, py = x[..., None] / dim_t, y[..., None] / dim_t
    px = torch.stack((px[..., 0::2].sin(), px[..., 1::2].cos()), dim=4).flatten(3)
    py = torch.stack((py[..., 0::2].sin(), py[..., 1::2].cos()), dim=4).flatten(3)
    return torch.cat((py, px), dim=3)


class KernelDetr(nn.Module):
    """`pred_logits, pred_boxes = net(images (B,3,H,W) fp32, mask (B,H,W) bool)` for an injected stock DETR `detector`."""

    D, HEADS = 256, 8

    def __init__(self, detector: nn.Module, with_backbone: bool = True):
        super().__init__()
        tr = detector.transformer
        dev = detector.query_embed.weight.device
        if dev.type != "cuda":
            raise ValueError("KernelDetr needs the detector on a CUDA device")
        if getattr(tr, "d_model", self.D) != self.D or getattr(tr, "nhead", self.HEADS) != self.HEADS:
            raise NotImplementedError("KernelDetr is built for d_model 256 / 8 heads")
        enc0 = tr.encoder.layers[0]
        if getattr(enc0, "normalize_before", False) or getattr(tr.encoder, "norm", None) is not None:
            raise NotImplementedError("KernelDetr implements the post-norm transformer (args.pre_norm = False)")
        self.device_ = dev
        self._keep = []
        bf = lambda t: self._hold(t.detach().to(torch.bfloat16).contiguous())
        f32 = lambda t: self._hold(t.detach().float().contiguous())
        D = self.D

        def attn(m):      # nn.MultiheadAttention -> packed [q; k] (512, 256), v, out-proj
            w, b = m.in_proj_weight, m.in_proj_bias
            return dict(wqk=bf(w[: 2 * D]), bqk=f32(b[: 2 * D]), wq=bf(w[:D]), bq=f32(b[:D]), wk=bf(w[D: 2 * D]), bk=f32(b[D: 2 * D]),
                        wv=bf(w[2 * D:]), bv=f32(b[2 * D:]), wo=bf(m.out_proj.weight), bo=f32(m.out_proj.bias))

        def ffn(l):
            return dict(w1=bf(l.linear1.weight), b1=f32(l.linear1.bias), w2=bf(l.linear2.weight), b2=f32(l.linear2.bias))

        norm = lambda n: (f32(n.weight), f32(n.bias))
        self.enc = [dict(sa=attn(l.self_attn), ff=ffn(l), n1=norm(l.norm1), n2=norm(l.norm2)) for l in tr.encoder.layers]
        self.dec = [dict(sa=attn(l.self_attn), ca=attn(l.multihead_attn), ff=ffn(l), n1=norm(l.norm1), n2=norm(l.norm2),
                         n3=norm(l.norm3)) for l in tr.decoder.layers]
        self.dec_norm = norm(tr.decoder.norm)
        self.w_in = bf(detector.input_proj.weight.reshape(D, -1))
        self.b_in = f32(detector.input_proj.bias)
        self.query_pos = f32(detector.query_embed.weight)
        self.num_queries = int(self.query_pos.shape[0])
        self.w_cls, self.b_cls = bf(detector.class_embed.weight), f32(detector.class_embed.bias)
        self.mlp = [(bf(l.weight), f32(l.bias)) for l in detector.bbox_embed.layers]
        self.backbone = KernelResNet50(detector.backbone[0].body) if with_backbone else None
        self._ws: Dict[tuple, dict] = {}

    def _hold(self, t):
        self._keep.append(t)
        return t

    # ---- per (batch, tokens, stream) scratch ----
    def _workspace(self, B: int, L: int) -> dict:
        key = (B, L, torch.cuda.current_stream(self.device_).cuda_stream)
        ws = self._ws.get(key)
        if ws is None:
            if len(self._ws) >= 8:
                self._ws.clear()
            dev, D, Q = self.device_, self.D, self.num_queries
            e = lambda r, c, dt=torch.bfloat16: torch.empty(r, c, device=dev, dtype=dt)
            n, m = B * L, B * Q
            ws = self._ws[key] = dict(
                x=e(n, D, torch.float32), xb=e(n, D), xpb=e(n, D), qk=e(n, 2 * D), v=e(n, D), att=e(n, D), delta=e(n, D), hdn=e(n, 2048),
                ck=e(n, D), cv=e(n, D),
                tgt=e(m, D, torch.float32), tb=e(m, D), tpb=e(m, D), dqk=e(m, 2 * D), dv=e(m, D), datt=e(m, D), ddelta=e(m, D),
                dhdn=e(m, 2048), cq=e(m, D), h1=e(m, D), h2=e(m, D),
                logits=e(m, int(self.w_cls.shape[0]), torch.float32), boxes=e(m, 4, torch.float32))
        return ws

    @staticmethod
    def _ln(x, delta, norm, pos, pos_rows, xb, xpb):
        _cabi.call("hoigen_add_layernorm256", x.data_ptr(), delta.data_ptr() if delta is not None else None,
                   norm[0].data_ptr() if norm is not None else None, norm[1].data_ptr() if norm is not None else None,
                   pos.data_ptr() if pos is not None else None, int(pos_rows), xb.data_ptr() if xb is not None else None,
                   xpb.data_ptr() if xpb is not None else None, int(x.shape[0]))

    def _attention(self, q, k, v, out, mask, B, lq, lk):
        _cabi.call("hoigen_attention_heads32", q.data_ptr(), q.stride(0), k.data_ptr(), k.stride(0), v.data_ptr(), v.stride(0),
                   out.data_ptr(), out.stride(0), mask.data_ptr() if mask is not None else None, B, lq, lk, self.HEADS,
                   1.0 / math.sqrt(self.D // self.HEADS))

    @torch.no_grad()
    def head(self, rows: torch.Tensor, key_mask: torch.Tensor, pos: torch.Tensor, B: int):
        """rows (B * L, 2048) bf16 backbone features (any token layout), key_mask (B, L) uint8 (1 = not a real pixel), pos (B * L, 256)
        fp32 -> pred_logits (B, Q, C + 1), pred_boxes (B, Q, 4) fp32 of the last decoder layer."""
        _cabi.init(rows.device)
        L, D, Q = rows.shape[0] // B, self.D, self.num_queries
        ws = self._workspace(B, L)
        G, relu = _cabi.gemm_bf16, _cabi.ACT_RELU
        x, xb, xpb, qk, v, att, delta, hdn = (ws[k] for k in ("x", "xb", "xpb", "qk", "v", "att", "delta", "hdn"))
        G(rows, self.w_in, bias=self.b_in, out_f32=x)                                        # input_proj (1x1 convolution)
        self._ln(x, None, None, pos, B * L, xb, xpb)
        for l in self.enc:                                                                   # transformer.py:134-150
            G(xpb, l["sa"]["wqk"], bias=l["sa"]["bqk"], out_bf16=qk)
            G(xb, l["sa"]["wv"], bias=l["sa"]["bv"], out_bf16=v)
            self._attention(qk[:, :D], qk[:, D:], v, att, key_mask, B, L, L)
            G(att, l["sa"]["wo"], bias=l["sa"]["bo"], out_bf16=delta)
            self._ln(x, delta, l["n1"], pos, B * L, xb, xpb)
            G(xb, l["ff"]["w1"], bias=l["ff"]["b1"], act=relu, out_bf16=hdn)
            G(hdn, l["ff"]["w2"], bias=l["ff"]["b2"], out_bf16=delta)
            self._ln(x, delta, l["n2"], pos, B * L, xb, xpb)
        # memory = x: xb = bf16(memory), xpb = bf16(memory + pos)
        tgt, tb, tpb, dqk, dv, datt, dd, dh, cq, ck, cv = (ws[k] for k in ("tgt", "tb", "tpb", "dqk", "dv", "datt", "ddelta", "dhdn",
                                                                             "cq", "ck", "cv"))
        tgt.zero_()
        self._ln(tgt, None, None, self.query_pos, Q, tb, tpb)
        for l in self.dec:                                                                   # transformer.py:187-209
            G(tpb, l["sa"]["wqk"], bias=l["sa"]["bqk"], out_bf16=dqk)
            G(tb, l["sa"]["wv"], bias=l["sa"]["bv"], out_bf16=dv)
            self._attention(dqk[:, :D], dqk[:, D:], dv, datt, None, B, Q, Q)
            G(datt, l["sa"]["wo"], bias=l["sa"]["bo"], out_bf16=dd)
            self._ln(tgt, dd, l["n1"], self.query_pos, Q, tb, tpb)
            G(tpb, l["ca"]["wq"], bias=l["ca"]["bq"], out_bf16=cq)
            G(xpb, l["ca"]["wk"], bias=l["ca"]["bk"], out_bf16=ck)
            G(xb, l["ca"]["wv"], bias=l["ca"]["bv"], out_bf16=cv)
            self._attention(cq, ck, cv, datt, key_mask, B, Q, L)
            G(datt, l["ca"]["wo"], bias=l["ca"]["bo"], out_bf16=dd)
            self._ln(tgt, dd, l["n2"], self.query_pos, Q, tb, tpb)
            G(tb, l["ff"]["w1"], bias=l["ff"]["b1"], act=relu, out_bf16=dh)
            G(dh, l["ff"]["w2"], bias=l["ff"]["b2"], out_bf16=dd)
            self._ln(tgt, dd, l["n3"], self.query_pos, Q, tb, tpb)
        self._ln(tgt, None, self.dec_norm, None, 0, tb, None)                                # decoder.norm (transformer.py:113-114)
        G(tb, self.w_cls, bias=self.b_cls, out_f32=ws["logits"])                             # class_embed
        G(tb, self.mlp[0][0], bias=self.mlp[0][1], act=relu, out_bf16=ws["h1"])              # bbox_embed: 3-layer MLP
        G(ws["h1"], self.mlp[1][0], bias=self.mlp[1][1], act=relu, out_bf16=ws["h2"])
        G(ws["h2"], self.mlp[2][0], bias=self.mlp[2][1], out_f32=ws["boxes"])
        return ws["logits"].view(B, Q, -1).clone(), ws["boxes"].sigmoid().view(B, Q, 4)

    @torch.no_grad()
    def head_from_features(self, src: torch.Tensor, mask: torch.Tensor):
        """src (B, 2048, h, w) fp32 backbone features, mask (B, h, w) bool -> (pred_logits, pred_boxes); tests / stock backbones."""
        B, C, h, w = src.shape
        rows = src.permute(0, 2, 3, 1).reshape(B * h * w, C).to(torch.bfloat16).contiguous()
        pos = sine_position_embedding(mask).reshape(B * h * w, self.D).contiguous()
        return self.head(rows, mask.reshape(B, h * w).to(torch.uint8).contiguous(), pos, B)

    @torch.no_grad()
    def forward(self, images: torch.Tensor, mask: torch.Tensor):
        """images (B, 3, H, W) fp32 padded batch, mask (B, H, W) bool (True = padding): what NestedTensor carries (U:1593)."""
        if self.backbone is None:
            raise ValueError("KernelDetr was built without its backbone; use head_from_features")
        B = images.shape[0]
        rows, (h, w) = self.backbone.features(images)
        m = F.interpolate(mask[None].float(), size=(h, w)).to(torch.bool)[0]                 # detr/models/backbone.py:77
        pos = F.pad(sine_position_embedding(m), (0, 0, 1, 1, 1, 1))                          # zeros on the halo ring
        key_mask = F.pad(m, (1, 1, 1, 1), value=True)                                        # halo pixels are padding keys
        L = (h + 2) * (w + 2)
        return self.head(rows, key_mask.reshape(B, L).to(torch.uint8).contiguous(), pos.reshape(B * L, self.D).contiguous(), B)
