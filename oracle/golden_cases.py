"""The golden-fixture cases shared by oracle/make_golden.py (which runs the UNMODIFIED reference on them and commits its
outputs as tests/golden/<name>.npz) and by the parity tests (CPU: oracle vs fixture; GPU: CUDA path vs fixture).
Test infrastructure: inputs are re-created from seeds, only reference OUTPUTS are stored.

`compact` cases are the BASELINE.json configurations at their benchmarked cache sizes; their fixtures hold the
detections, logits and global feature only (no 400 KB/image token maps)."""
from __future__ import annotations

import torch

from hoigen_b200 import synthetic as S

CASES = {
    "hico117_b2": dict(num_classes=117, dataset="hicodet", B=2, n_h=8, n_o=8, N=256, ragged=False, boxes="grid"),
    "hico117_ragged_b3": dict(num_classes=117, dataset="hicodet", B=3, n_h=6, n_o=7, N=234, ragged=True, boxes="grid"),
    "hico117_oob_b1": dict(num_classes=117, dataset="hicodet", B=1, n_h=4, n_o=5, N=128, ragged=False, boxes="oob"),
    "vcoco24_b2": dict(num_classes=24, dataset="vcoco", B=2, n_h=16, n_o=16, N=96, ragged=False, boxes="grid",
                       args=dict(max_instances=16, cache=True, eval=False)),
    # BASELINE configs[4]'s classifier: 600 HOI triplets (`generate_feature=False`, class_corr = object -> interaction)
    "hico600_b2": dict(num_classes=600, dataset="hicodet", B=2, n_h=8, n_o=8, N=600, ragged=False, boxes="grid",
                       args=dict(generate_feature=False)),
    # ---- the BASELINE configurations at their benchmarked cache sizes ------------------------------------------------
    # configs[0]/[1]: HICO-117, 8h+8o = 120 pairs, 4096 x 512 caches
    "hico117_n4096_b4": dict(num_classes=117, dataset="hicodet", B=4, n_h=8, n_o=8, N=4096, ragged=False, boxes="grid",
                             compact=True),
    # configs[2]: uc0 zero-shot, 16384 x 512 caches ending in generator-synthesised rows of the 120 held-out HOI classes
    "hico117_uc0_n16384_b2": dict(num_classes=117, dataset="hicodet", B=2, n_h=8, n_o=8, N=16384, ragged=False, boxes="grid",
                                  compact=True, head="uc0", args=dict(zs=True, zs_type="uc0")),
    # configs[3]: V-COCO 24 actions, 16h+16o = 496 pairs, 4096 x 512 caches
    "vcoco24_n4096_b2": dict(num_classes=24, dataset="vcoco", B=2, n_h=16, n_o=16, N=4096, ragged=False, boxes="grid",
                             compact=True, args=dict(max_instances=16, cache=True, eval=False)),
    # configs[4]: 600-triplet HICO classifier, 4096 x 512 caches
    "hico600_n4096_b2": dict(num_classes=600, dataset="hicodet", B=2, n_h=8, n_o=8, N=4096, ragged=False, boxes="grid",
                             compact=True, args=dict(generate_feature=False)),
}

_uc0_cache = {}


def head_for(case) -> S.HeadState:
    over = case.get("args", {})
    if case.get("head") == "uc0":
        if case["N"] not in _uc0_cache:          # ~10 s of CPU text-encoder work: once per process
            _uc0_cache[case["N"]] = S.make_head_state_uc0(case["N"], seed=2, gen_seed=3)
        return _uc0_cache[case["N"]]
    return S.make_head_state(case["num_classes"], case["N"], seed=2, max_instances=over.get("max_instances", 15))


def props_for(case):
    props = S.make_region_props(case["B"], case["n_h"], case["n_o"], ragged=case["ragged"])
    if case["boxes"] == "oob":
        # boxes partly / mostly outside the 224^2 image: exercises the `< -1 / > size` zeroing rule and clamping
        g = torch.Generator().manual_seed(4242)
        for p in props:
            n = p["boxes"].shape[0]
            shift = (torch.rand(n, 2, generator=g) - 0.5) * 260.0
            p["boxes"] = p["boxes"] + torch.cat([shift, shift], dim=1)
            p["boxes"][0] = torch.tensor([-40.0, -30.0, 20.0, 260.0])
            p["boxes"][-1] = torch.tensor([100.0, 180.0, 330.0, 300.0])
    return props


def inputs_for(case):
    """-> (images (B,3,224,224), region proposals, dino features (B,2048))"""
    return S.make_images(case["B"], seed=1), props_for(case), S.make_dino_features(case["B"])
