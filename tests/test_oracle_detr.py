"""CPU: the DETR oracle (oracle/detr_ref.py) reproduces the outputs of the UNMODIFIED reference DETR classes that
oracle/make_golden_detr.py stored (tests/golden/detr_head.npz), from weights re-created out of the seed."""
import json

import numpy as np
import torch


def test_detr_oracle_reproduces_reference_golden():
    from oracle import detr_ref as D
    gold = np.load("tests/golden/detr_head.npz")
    det = D.DetrRef(num_classes=int(gold["classes"])).eval()
    D.seeded_state(det, int(gold["seed"]))
    logits, boxes = det.forward_features(torch.from_numpy(gold["src"].astype(np.float32)), torch.from_numpy(gold["mask"]))
    assert (logits - torch.from_numpy(gold["logits"])).abs().max().item() < 1e-4
    assert (boxes - torch.from_numpy(gold["boxes"])).abs().max().item() < 1e-5
    pin = json.load(open("tests/golden/PINNING.json"))["detr_head"]
    assert pin["logits_max_abs"] < 1e-4 and pin["boxes_max_abs"] < 1e-5 and pin["pos_embedding_max_abs"] < 1e-5


def test_detr_oracle_transformer_forward_equals_feature_path():
    """The nn.Module-style entry the stock path calls (transformer(src, mask, query_embed, pos), U:1596) and the fused
    `forward_features` are the same arithmetic."""
    from oracle import detr_ref as D
    torch.manual_seed(0)
    det = D.DetrRef().eval()
    D.seeded_state(det, 3)
    src = torch.randn(2, 2048, 3, 4).clamp_min(0)
    mask = torch.zeros(2, 3, 4, dtype=torch.bool)
    mask[1, :, 3:] = True
    with torch.no_grad():
        hs, _ = det.transformer(det.input_proj(src), mask, det.query_embed.weight, D.sine_position_embedding(mask))
        a = det.class_embed(hs)[-1], det.bbox_embed(hs).sigmoid()[-1]
        b = det.forward_features(src, mask)
    assert torch.allclose(a[0], b[0], atol=1e-6) and torch.allclose(a[1], b[1], atol=1e-6)
