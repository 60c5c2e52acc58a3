"""TEST INFRASTRUCTURE — CPU / torch restatement of the DETR detector behind UPT.forward's proposal stage:
position encoding, input projection, 6 + 6-layer post-norm transformer, class / box heads.  Only tests/ may import this.

Follows the reference file by file (nothing is copied; the attribute NAMES are the reference's so that its state dict loads):
  * detr/models/position_encoding.py:12-50   PositionEmbeddingSine(128, normalize=True)   -> `sine_position_embedding`
  * detr/models/transformer.py:46-60         flatten, zeros tgt, encoder, decoder          -> `DetrRef.forward_features`
  * detr/models/transformer.py:127-150       encoder layer, forward_post                   -> `EncoderLayer`
  * detr/models/transformer.py:187-209       decoder layer, forward_post                   -> `DecoderLayer`
  * detr/models/transformer.py:100-121       decoder norm on the returned activations      -> `DetrRef.forward_features`
  * detr/models/detr.py:37-40, 62-68, 293-305  input_proj, query_embed, class_embed, bbox_embed (3-layer MLP), sigmoid
  * call site: upt_tip_cache_model_free_finetune_distill3.py:1594-1599

Pinned against the UNMODIFIED reference classes by oracle/make_golden_detr.py (same seeded state dict, CPU): see
tests/golden/PINNING.json ("detr_head") and tests/golden/detr_head.npz.
"""
from __future__ import annotations

import math

import torch
from torch import nn
import torch.nn.functional as F


def sine_position_embedding(mask: torch.Tensor, num_pos_feats: int = 128, temperature: float = 10000.0) -> torch.Tensor:
    """mask (B, h, w) bool, True = padding -> (B, 2 * num_pos_feats, h, w) fp32  (position_encoding.py:28-50, normalize=True)."""
    not_mask = ~mask
    y_embed = not_mask.cumsum(1, dtype=torch.float32)
    x_embed = not_mask.cumsum(2, dtype=torch.float32)
    eps, scale = 1e-6, 2 * math.pi
    y_embed = y_embed / (y_embed[:, -1:, :] + eps) * scale
    x_embed = x_embed / (x_embed[:, :, -1:] + eps) * scale
    dim_t = torch.arange(num_pos_feats, dtype=torch.float32, device=mask.device)
    dim_t = temperature ** (2 * (dim_t // 2) / num_pos_feats)
    px, py = x_embed[..., None] / dim_t, y_embed[..., None] / dim_t
    px = torch.stack((px[..., 0::2].sin(), px[..., 1::2].cos()), dim=4).flatten(3)
    py = torch.stack((py[..., 0::2].sin(), py[..., 1::2].cos()), dim=4).flatten(3)
    return torch.cat((py, px), dim=3).permute(0, 3, 1, 2)


class EncoderLayer(nn.Module):
    def __init__(self, d=256, heads=8, ff=2048):
        super().__init__()
        self.self_attn = nn.MultiheadAttention(d, heads, dropout=0.1)
        self.linear1, self.linear2 = nn.Linear(d, ff), nn.Linear(ff, d)
        self.norm1, self.norm2 = nn.LayerNorm(d), nn.LayerNorm(d)

    def forward(self, src, key_padding_mask, pos):
        q = k = src + pos
        src = self.norm1(src + self.self_attn(q, k, value=src, key_padding_mask=key_padding_mask)[0])
        return self.norm2(src + self.linear2(F.relu(self.linear1(src))))


class DecoderLayer(nn.Module):
    def __init__(self, d=256, heads=8, ff=2048):
        super().__init__()
        self.self_attn = nn.MultiheadAttention(d, heads, dropout=0.1)
        self.multihead_attn = nn.MultiheadAttention(d, heads, dropout=0.1)
        self.linear1, self.linear2 = nn.Linear(d, ff), nn.Linear(ff, d)
        self.norm1, self.norm2, self.norm3 = nn.LayerNorm(d), nn.LayerNorm(d), nn.LayerNorm(d)

    def forward(self, tgt, memory, memory_key_padding_mask, pos, query_pos):
        q = k = tgt + query_pos
        tgt = self.norm1(tgt + self.self_attn(q, k, value=tgt)[0])
        tgt = self.norm2(tgt + self.multihead_attn(query=tgt + query_pos, key=memory + pos, value=memory,
                                                   key_padding_mask=memory_key_padding_mask)[0])
        return self.norm3(tgt + self.linear2(F.relu(self.linear1(tgt))))


class _Stack(nn.Module):
    def __init__(self, layers, norm=None):
        super().__init__()
        self.layers = nn.ModuleList(layers)
        if norm is not None:
            self.norm = norm


class _Transformer(nn.Module):
    def __init__(self, d=256, heads=8, ff=2048, enc=6, dec=6):
        super().__init__()
        self.encoder = _Stack([EncoderLayer(d, heads, ff) for _ in range(enc)])
        self.decoder = _Stack([DecoderLayer(d, heads, ff) for _ in range(dec)], nn.LayerNorm(d))

    def forward(self, src, mask, query_embed, pos_embed):
        """transformer.py:46-60 (only the last decoder layer is returned, stacked as one: all U:1604 reads)."""
        bs, c, h, w = src.shape
        x = src.flatten(2).permute(2, 0, 1)
        pos = pos_embed.flatten(2).permute(2, 0, 1)
        qpos = query_embed.unsqueeze(1).repeat(1, bs, 1)
        kpm = mask.flatten(1)
        for layer in self.encoder.layers:
            x = layer(x, kpm, pos)
        tgt = torch.zeros_like(qpos)
        for layer in self.decoder.layers:
            tgt = layer(tgt, x, kpm, pos, qpos)
        return self.decoder.norm(tgt).transpose(0, 1)[None], x.permute(1, 2, 0).view(bs, c, h, w)


class _MLP(nn.Module):
    def __init__(self, d):
        super().__init__()
        self.layers = nn.ModuleList([nn.Linear(d, d), nn.Linear(d, d), nn.Linear(d, 4)])

    def forward(self, x):
        for i, lin in enumerate(self.layers):
            x = F.relu(lin(x)) if i < 2 else lin(x)
        return x


class DetrRef(nn.Module):
    """The part of the DETR detector UPT.forward evaluates after the backbone (U:1595-1599), last decoder layer only."""

    def __init__(self, num_classes=91, num_queries=100, d=256, channels=2048):
        super().__init__()
        self.transformer = _Transformer(d)
        self.class_embed = nn.Linear(d, num_classes + 1)
        self.bbox_embed = _MLP(d)
        self.query_embed = nn.Embedding(num_queries, d)
        self.input_proj = nn.Conv2d(channels, d, kernel_size=1)

    @torch.no_grad()
    def forward_features(self, src: torch.Tensor, mask: torch.Tensor):
        """src (B, 2048, h, w) backbone features, mask (B, h, w) bool -> pred_logits (B, Q, C + 1), pred_boxes (B, Q, 4)."""
        pos = sine_position_embedding(mask)
        x = self.input_proj(src)
        bs = x.shape[0]
        x = x.flatten(2).permute(2, 0, 1)
        pos = pos.flatten(2).permute(2, 0, 1)
        qpos = self.query_embed.weight.unsqueeze(1).repeat(1, bs, 1)
        kpm = mask.flatten(1)
        for layer in self.transformer.encoder.layers:
            x = layer(x, kpm, pos)
        tgt = torch.zeros_like(qpos)
        for layer in self.transformer.decoder.layers:
            tgt = layer(tgt, x, kpm, pos, qpos)
        hs = self.transformer.decoder.norm(tgt).transpose(0, 1)
        boxes = hs
        for i, lin in enumerate(self.bbox_embed.layers):
            boxes = F.relu(lin(boxes)) if i < 2 else lin(boxes)
        return self.class_embed(hs), boxes.sigmoid()


def seeded_state(module: nn.Module, seed: int) -> None:
    """Fill every parameter of `module` from a CPU generator in state-dict order (the same numbers on any machine): weights
    ~ N(0, 1) / sqrt(fan_in), biases ~ 0.1 N(0, 1), LayerNorm weights 1 + 0.1 N(0, 1), the query embedding ~ N(0, 1)."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    with torch.no_grad():
        for name, p in module.state_dict().items():
            r = torch.randn(p.shape, generator=g)
            if "norm" in name and name.endswith("weight"):
                v = 1.0 + 0.1 * r
            elif name.endswith("bias"):
                v = 0.1 * r
            elif name.startswith("query_embed"):
                v = r
            else:
                fan_in = p[0].numel() if p.dim() > 1 else p.numel()
                v = r / math.sqrt(fan_in)
            p.copy_(v.to(p.dtype))
