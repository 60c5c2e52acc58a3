"""Pin oracle/detr_ref.py against the UNMODIFIED reference DETR classes (builder container only: needs /root/reference).

    python oracle/make_golden_detr.py     # writes tests/golden/detr_head.npz, merges the record into tests/golden/PINNING.json

Reference leg: detr/models/transformer.py::Transformer (6 + 6 layers, post-norm, return_intermediate_dec=True as
build_transformer sets it), detr/models/position_encoding.py::PositionEmbeddingSine(128, normalize=True), detr/models/detr.py::MLP,
nn.Conv2d input_proj / nn.Linear class_embed / nn.Embedding query_embed as DETR.__init__ creates them, evaluated exactly as
U:1595-1599 does.  Both legs get the same state dict (oracle.detr_ref.seeded_state).  Only the reference's OUTPUTS and the small
seeded inputs are stored; the 18 M weights are re-created from the seed by the tests.
"""
from __future__ import annotations

import json
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, "/root/reference/detr")

from oracle import detr_ref as D  # noqa: E402

SEED, B, H, W, CLASSES = 17, 2, 5, 7, 91


def inputs():
    g = torch.Generator().manual_seed(SEED + 1)
    src = (torch.randn(B, 2048, H, W, generator=g).clamp_min(0) * 0.5).half().float()   # ReLU-like backbone features, fp16-exact (stored as fp16)
    mask = torch.zeros(B, H, W, dtype=torch.bool)
    mask[1, :, 5:] = True                                                     # image 1 is narrower: right columns are padding
    mask[1, 4:, :] = True
    return src, mask


def main():
    from models.transformer import Transformer
    from models.position_encoding import PositionEmbeddingSine
    from models.detr import MLP
    from torch import nn
    torch.manual_seed(0)
    mine = D.DetrRef(num_classes=CLASSES).eval()
    D.seeded_state(mine, SEED)
    sd = mine.state_dict()
    ref_tr = Transformer(d_model=256, dropout=0.1, nhead=8, dim_feedforward=2048, num_encoder_layers=6, num_decoder_layers=6,
                         normalize_before=False, return_intermediate_dec=True).eval()
    ref_tr.load_state_dict({k[len("transformer."):]: v for k, v in sd.items() if k.startswith("transformer.")}, strict=True)
    ref_pos = PositionEmbeddingSine(128, normalize=True)
    ref_mlp = MLP(256, 256, 4, 3).eval()
    ref_mlp.load_state_dict({k[len("bbox_embed."):]: v for k, v in sd.items() if k.startswith("bbox_embed.")}, strict=True)
    input_proj, class_embed, query_embed = nn.Conv2d(2048, 256, 1), nn.Linear(256, CLASSES + 1), nn.Embedding(100, 256)
    input_proj.load_state_dict({"weight": sd["input_proj.weight"], "bias": sd["input_proj.bias"]})
    class_embed.load_state_dict({"weight": sd["class_embed.weight"], "bias": sd["class_embed.bias"]})
    query_embed.load_state_dict({"weight": sd["query_embed.weight"]})
    src, mask = inputs()

    class NT:                                                                  # what PositionEmbeddingSine.forward reads
        tensors, mask = src, None
    NT.mask = mask
    with torch.no_grad():
        pos = ref_pos(NT)
        hs = ref_tr(input_proj(src), mask, query_embed.weight, pos)[0]          # U:1596
        ref_logits = class_embed(hs)[-1]                                        # U:1597, U:1604 (last decoder layer)
        ref_boxes = ref_mlp(hs).sigmoid()[-1]
        my_logits, my_boxes = mine.forward_features(src, mask)
        pos_err = (D.sine_position_embedding(mask) - pos).abs().max().item()
    rec = {"case": f"seed {SEED}, B={B}, feature map {H}x{W} with padding, {CLASSES}+1 classes",
           "pos_embedding_max_abs": pos_err,
           "logits_max_abs": (my_logits - ref_logits).abs().max().item(),
           "boxes_max_abs": (my_boxes - ref_boxes).abs().max().item(),
           "logits_abs_max": ref_logits.abs().max().item()}
    print(json.dumps(rec, indent=1))
    assert rec["logits_max_abs"] < 1e-4 and rec["boxes_max_abs"] < 1e-5 and pos_err < 1e-5, rec
    out = ROOT / "tests" / "golden"
    np.savez_compressed(out / "detr_head.npz", src=src.numpy().astype(np.float16), mask=mask.numpy(),
                        logits=ref_logits.numpy(), boxes=ref_boxes.numpy(), seed=np.int64(SEED), classes=np.int64(CLASSES))
    pin_path = out / "PINNING.json"
    pin = json.loads(pin_path.read_text()) if pin_path.exists() else {}
    pin["detr_head"] = rec
    pin_path.write_text(json.dumps(pin, indent=1) + "\n")


if __name__ == "__main__":
    main()
