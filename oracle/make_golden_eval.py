"""Pin oracle/eval_ref.py against the UNMODIFIED reference (builder container only: needs /root/reference).

    python oracle/make_golden_eval.py     # writes tests/golden/assoc_*.npz, prints the comparison

The reference leg is the loop body of `CustomisedDLE.test_hico` (utils_tip_cache_and_union_finetune.py:375-407) run with
the real `pocket.utils.BoxPairAssociation`, the real `conversion[objects, verbs]` lookup and the real
`UPT.recover_boxes` arithmetic (detr `box_cxcywh_to_xyxy`), on seeded synthetic detections / targets
(oracle/eval_ref.py::synthetic_detections / make_targets).  Only the reference's OUTPUTS (interactions, labels) are
stored; the inputs are re-created from the seeds by the tests.
"""
from __future__ import annotations

import json
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, "/root/reference/pocket")
sys.path.insert(0, "/root/reference/detr")

from oracle import eval_ref as E  # noqa: E402

CASES = {"assoc_b6": dict(batch=6, seed=11, n_h=4, n_o=4), "assoc_b3_dense": dict(batch=3, seed=12, n_h=6, n_o=2)}


def reference_loop(outputs, targets, conversion_list):
    """T:352-407 with the reference's own objects (no restatement on this side)."""
    from pocket.utils import BoxPairAssociation
    from util import box_ops                                            # detr/util/box_ops.py, as U:1270 uses it
    associate = BoxPairAssociation(min_iou=0.5)
    conversion = torch.from_numpy(np.asarray(conversion_list, dtype=float))

    def recover_boxes(boxes, size):                                      # U:1269-1274 verbatim arithmetic
        boxes = box_ops.box_cxcywh_to_xyxy(boxes)
        h, w = size
        return boxes * torch.stack([w, h, w, h])

    res = []
    for output, target in zip(outputs, targets):
        boxes = output["boxes"]
        boxes_h, boxes_o = boxes[output["pairing"]].unbind(0)
        objects, scores, verbs = output["objects"], output["scores"], output["labels"]
        interactions = conversion[objects, verbs]
        gt_bx_h = recover_boxes(target["boxes_h"], target["size"])
        gt_bx_o = recover_boxes(target["boxes_o"], target["size"])
        labels = torch.zeros_like(scores)
        for hoi_idx in interactions.unique():
            gt_idx = torch.nonzero(target["hoi"] == hoi_idx).squeeze(1)
            det_idx = torch.nonzero(interactions == hoi_idx).squeeze(1)
            if len(gt_idx):
                labels[det_idx] = associate((gt_bx_h[gt_idx].view(-1, 4), gt_bx_o[gt_idx].view(-1, 4)),
                                            (boxes_h[det_idx].view(-1, 4), boxes_o[det_idx].view(-1, 4)),
                                            scores[det_idx].view(-1))
        res.append((scores, interactions, labels))
    return res


def main():
    from hoigen_b200 import synthetic as S
    tables = json.load(open(Path(S.__file__).parent / "data" / "object_tables.json"))
    onv = [[(v if v >= 0 else None) for v in row] for row in tables["hico_object_n_verb_to_interaction"]]
    conv = E.conversion_table(onv)
    for name, c in CASES.items():
        dets = E.synthetic_detections(c["batch"], c["seed"], c["n_h"], c["n_o"])
        tgts = E.make_targets(dets, conv, seed=c["seed"] + 1000)
        ref = reference_loop(dets, tgts, onv)
        mine = E.associate_batch(dets, tgts, conv)
        out = {}
        tp = 0
        for b, ((rs, ri, rl), (ms, mi, ml)) in enumerate(zip(ref, mine)):
            assert torch.equal(ri.nan_to_num(-1), mi.nan_to_num(-1)), (name, b, "interactions")
            assert torch.equal(rl, ml), (name, b, "labels")
            out[f"interactions_{b}"] = ri.numpy()
            out[f"labels_{b}"] = rl.numpy()
            tp += int(rl.sum())
        np.savez_compressed(ROOT / "tests" / "golden" / f"{name}.npz", **out)
        print(f"{name}: oracle == reference on {len(ref)} images, {sum(int(r[0].numel()) for r in ref)} detections, {tp} true positives")


def main_ap():
    """DetectionAPMeter('11P', precision 64, nproc 1) of the reference on a seeded sweep -> tests/golden/ap11_*.npz."""
    from pocket.utils import DetectionAPMeter
    for name, seed, with_gt in (("ap11_with_num_gt", 21, True), ("ap11_no_num_gt", 22, False)):
        stream, num_gt = E.synthetic_meter_stream(seed)
        meter = DetectionAPMeter(600, nproc=1, num_gt=num_gt if with_gt else None, algorithm="11P")
        for sc, pr, lb in stream:
            meter.append(sc, pr, lb)
        ap = meter.eval()
        sc_c, lb_c = E.group_by_class(stream, 600)
        o_ap, o_rec = E.ap_11point(sc_c, lb_c, num_gt if with_gt else None)
        assert torch.equal(ap, o_ap), (name, (ap - o_ap).abs().max())
        assert torch.equal(meter.max_rec.nan_to_num(-1), o_rec.nan_to_num(-1)), name
        np.savez_compressed(ROOT / "tests" / "golden" / f"{name}.npz", ap=ap.numpy(), max_rec=meter.max_rec.numpy())
        print(f"{name}: oracle == reference bit for bit over 600 classes, mAP {ap.mean().item():.6f}")


if __name__ == "__main__":
    main()
    main_ap()
