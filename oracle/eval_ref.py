"""CPU restatement of the reference's evaluation-side association — TEST INFRASTRUCTURE ONLY.

Only tests/ and tools that check the CUDA path may import this module; the product path (hoigen_b200/evaluate.py ->
hoigen_associate_pairs) never does.

Restates, does not copy:
  * `CustomisedDLE.test_hico`, utils_tip_cache_and_union_finetune.py:348-411 (T) — per image: HOI id of every detection
    through the `object_n_verb_to_interaction` table, then for every HOI id present among the detections that also has
    ground truth, associate the detections of that id with the ground-truth pairs of that id;
  * `BoxAssociation.__call__` / `BoxPairAssociation._iou`, pocket/pocket/utils/association.py:51-125 (A) — pair IoU =
    min(IoU(human boxes), IoU(object boxes)); every detection goes to the ground truth with the highest IoU (first
    index on ties); a ground truth's true positive is its highest-scoring detection among those with IoU > min_iou
    (first index on ties);
  * torchvision.ops.box_iou 0.26.0 (pocket.ops.box_iou with encoding='coord', pocket/pocket/ops/boxes.py:133-134):
    area = (x2-x1)*(y2-y1); inter = clamp(min(x2)-max(x1),0) * clamp(min(y2)-max(y1),0);
    iou = inter / ((area1 + area2) - inter), every operation rounded to fp32 separately (no FMA);
  * `UPT.recover_boxes`, upt_tip_cache_model_free_finetune_distill3.py:1269-1274 — cxcywh in [0,1] -> xyxy pixels.

Pinned by oracle/make_golden_eval.py against the UNMODIFIED pocket.utils.BoxPairAssociation and the reference's loop
(tests/golden/assoc_*.npz) and by the reference's own known-answer test pocket/test/association.py
(tests/test_eval_oracle.py).
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch


def recover_boxes(boxes_cxcywh: torch.Tensor, size: torch.Tensor) -> torch.Tensor:
    """U:1269-1274: (cx,cy,w,h) normalised -> (x1,y1,x2,y2) in pixels of an image of `size` = (h, w)."""
    cx, cy, w, h = boxes_cxcywh.unbind(-1)
    xyxy = torch.stack([cx - 0.5 * w, cy - 0.5 * h, cx + 0.5 * w, cy + 0.5 * h], dim=-1)
    ih, iw = size
    return xyxy * torch.stack([iw, ih, iw, ih]).to(xyxy.dtype)


def box_iou(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    """(N,4) x (M,4) -> (N,M), fp32, operation order of torchvision.ops.box_iou."""
    a, b = a.float(), b.float()
    area_a = (a[:, 2] - a[:, 0]) * (a[:, 3] - a[:, 1])
    area_b = (b[:, 2] - b[:, 0]) * (b[:, 3] - b[:, 1])
    lt = torch.maximum(a[:, None, :2], b[None, :, :2])
    rb = torch.minimum(a[:, None, 2:], b[None, :, 2:])
    wh = (rb - lt).clamp(min=0)
    inter = wh[..., 0] * wh[..., 1]
    union = area_a[:, None] + area_b[None, :] - inter
    return inter / union


def box_association(gt_iou: torch.Tensor, scores: Optional[torch.Tensor], min_iou: float) -> torch.Tensor:
    """A:51-95 given the (N_gt, M_det) IoU matrix: 0/1 label per detection."""
    n_gt, m = gt_iou.shape
    max_iou, max_idx = gt_iou.max(0)                      # first index on ties
    if scores is None:
        scores = max_iou
    labels = torch.zeros_like(scores)
    for g in range(n_gt):
        cand = torch.nonzero((max_idx == g) & (max_iou > min_iou)).squeeze(1)
        if len(cand) == 0:
            continue
        labels[cand[scores[cand].argmax()]] = 1           # first index on ties
    return labels


def pair_association(gt_h, gt_o, det_h, det_o, scores, min_iou: float = 0.5) -> torch.Tensor:
    """BoxPairAssociation (A:97-125): pair IoU = min of the human-box and object-box IoUs."""
    iou = torch.minimum(box_iou(gt_h, det_h), box_iou(gt_o, det_o))
    return box_association(iou, scores, min_iou)


def conversion_table(object_n_verb_to_interaction: Sequence[Sequence[Optional[int]]]) -> torch.Tensor:
    """T:356-358: float table with NaN where the (object, verb) combination is not an HOI class."""
    return torch.from_numpy(np.asarray([[np.nan if (v is None or v < 0) else float(v) for v in row]
                                        for row in object_n_verb_to_interaction], dtype=float))


def associate_image(output: dict, target: dict, conversion: Optional[torch.Tensor], min_iou: float = 0.5
                    ) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """T:375-407 for one image: -> (scores, interactions, labels) exactly as handed to `meter.append`.

    output: boxes (n,4), pairing (2,M), objects (M,), labels (M,) = verbs, scores (M,)
    target: boxes_h, boxes_o (G,4) normalised cxcywh, hoi (G,), size (2,) = (h, w)
    conversion None <=> the model's classes already are HOI ids (T:389-390)."""
    boxes = output["boxes"]
    boxes_h, boxes_o = boxes[output["pairing"]].unbind(0)
    objects, verbs, scores = output["objects"], output["labels"], output["scores"]
    interactions = conversion[objects, verbs] if conversion is not None else verbs
    gt_h = recover_boxes(target["boxes_h"], target["size"])
    gt_o = recover_boxes(target["boxes_o"], target["size"])
    labels = torch.zeros_like(scores)
    for hoi in interactions.unique():
        gt_idx = torch.nonzero(target["hoi"] == hoi).squeeze(1)
        det_idx = torch.nonzero(interactions == hoi).squeeze(1)
        if len(gt_idx):
            labels[det_idx] = pair_association(gt_h[gt_idx].view(-1, 4), gt_o[gt_idx].view(-1, 4),
                                               boxes_h[det_idx].view(-1, 4), boxes_o[det_idx].view(-1, 4),
                                               scores[det_idx].view(-1), min_iou)
    return scores, interactions, labels


def associate_batch(outputs: Sequence[dict], targets: Sequence[dict], conversion: Optional[torch.Tensor],
                    min_iou: float = 0.5) -> List[Tuple[torch.Tensor, torch.Tensor, torch.Tensor]]:
    return [associate_image(o, t, conversion, min_iou) for o, t in zip(outputs, targets)]


def make_targets(dets: Sequence[dict], conversion: torch.Tensor, seed: int, size=(224, 224), per_image: int = 6):
    """Synthetic ground truth for detections `dets` (test helper, shared by the golden generator and the GPU tests):
    per image a few ground-truth pairs copied from detected pairs with a small jitter (so IoU > 0.5 matches exist,
    including several detections competing for one ground truth and duplicated ground truths of one HOI id), a few with
    a large jitter (IoU < 0.5) and a few HOI ids that no detection has.  Boxes are returned in the reference's target
    format: normalised cxcywh, `size` = (h, w)."""
    g = torch.Generator().manual_seed(seed)
    ih, iw = size
    out = []
    for d in dets:
        m = int(d["scores"].numel())
        if m == 0:
            out.append(dict(boxes_h=torch.zeros(0, 4), boxes_o=torch.zeros(0, 4), hoi=torch.zeros(0, dtype=torch.int64),
                            size=torch.tensor([ih, iw], dtype=torch.float32)))
            continue
        boxes = d["boxes"].float().cpu()
        bh, bo = boxes[d["pairing"].cpu()].unbind(0)
        inter = conversion[d["objects"].cpu(), d["labels"].cpu()]
        pick = torch.randint(0, m, (per_image,), generator=g)
        pick = torch.cat([pick, pick[:2]])                                   # duplicated ground truths of one HOI id
        jit = torch.cat([torch.full((per_image - 2,), 3.0), torch.full((2,), 40.0), torch.full((2,), 2.0)])
        noise = (torch.rand(len(pick), 4, generator=g) - 0.5) * jit[:, None]
        gh = bh[pick] + noise
        go = bo[pick] + noise.flip(0)
        hoi = inter[pick].to(torch.int64)
        extra = torch.randint(0, 600, (2,), generator=g)                     # ids (probably) absent from the detections
        gh = torch.cat([gh, bh[:2]])
        go = torch.cat([go, bo[:2]])
        hoi = torch.cat([hoi, extra])

        def to_cxcywh(b):
            x1, y1, x2, y2 = b.unbind(-1)
            return torch.stack([(x1 + x2) / 2 / iw, (y1 + y2) / 2 / ih, (x2 - x1) / iw, (y2 - y1) / ih], dim=-1)
        out.append(dict(boxes_h=to_cxcywh(gh), boxes_o=to_cxcywh(go), hoi=hoi,
                        size=torch.tensor([ih, iw], dtype=torch.float32)))
    return out


def synthetic_detections(batch: int, seed: int, n_h: int = 4, n_o: int = 4, tie_every: int = 7) -> List[dict]:
    """Detections in the detector's output format without running the model (test helper): every (human, other box)
    pair of hoigen_b200.synthetic boxes, one triplet per verb the object class allows (HICO table), scores in (0,1)
    with deliberate exact ties every `tie_every`-th triplet (exercises the first-index tie rules)."""
    import json
    from pathlib import Path
    from hoigen_b200 import synthetic as S
    tables = json.load(open(Path(S.__file__).parent / "data" / "object_tables.json"))
    o2v = tables["hico_object_to_verb"]
    g = torch.Generator().manual_seed(seed)
    dets = []
    for b in range(batch):
        p = S.make_boxes(seed * 131 + b, n_h, n_o)
        # near-duplicate boxes (a second human ~ human 0, second copies of the first two objects): several detections
        # then overlap one ground-truth pair with IoU > 0.5 and compete for it by score
        jit = (torch.rand(3, 4, generator=g) - 0.5) * 3.0
        hb, ob = p["boxes"][:n_h], p["boxes"][n_h:]
        hl, ol = p["labels"][:n_h], p["labels"][n_h:]
        p = dict(boxes=torch.cat([hb, hb[:1] + jit[:1], ob, ob[:2] + jit[1:]]),
                 labels=torch.cat([hl, hl[:1], ol, ol[:2]]))
        n_hh = n_h + 1
        n = p["boxes"].shape[0]
        xs, ys, vs = [], [], []
        for x in range(n_hh):
            for y in range(n):
                if y == x:
                    continue
                for v in sorted(set(o2v[int(p["labels"][y])])):
                    xs.append(x); ys.append(y); vs.append(v)
        m = len(xs)
        scores = torch.rand(m, generator=g) * 0.98 + 0.01
        scores[tie_every::tie_every] = scores[0]
        dets.append(dict(boxes=p["boxes"], pairing=torch.tensor([xs, ys], dtype=torch.int64),
                         objects=p["labels"][torch.tensor(ys, dtype=torch.int64)] if m else torch.zeros(0, dtype=torch.int64),
                         labels=torch.tensor(vs, dtype=torch.int64), scores=scores,
                         size=torch.tensor([224, 224], dtype=torch.int64)))
    return dets


def ap_11point(scores: Sequence[torch.Tensor], labels: Sequence[torch.Tensor], num_gt=None):
    """DetectionAPMeter.compute_ap with algorithm '11P', precision 64 (pocket/pocket/utils/meters.py:493-583 and
    255-270): per class, order by score (descending; STABLE here, the reference's order among equal scores is
    unspecified), tp / fp cumulative sums, prec = tp / (tp + fp), rec = tp / num_gt (or / collected positives),
    ap = sum over t in linspace(0, 1, 11) of max{prec : rec >= t} / 11.  -> (ap, max_rec) fp64."""
    k = len(scores)
    ap = torch.zeros(k, dtype=torch.float64)
    max_rec = torch.zeros(k, dtype=torch.float64)
    thr = torch.linspace(0, 1, 11, dtype=torch.float64)
    for c in range(k):
        out, lab = scores[c].double(), labels[c].double()
        if not (len(out) and len(lab)):
            continue
        order = torch.argsort(out, descending=True, stable=True)
        tp = lab[order].cumsum(0)
        fp = (1 - lab[order]).cumsum(0)
        prec = tp / (tp + fp)
        denom = lab.sum().item() if (num_gt is None or num_gt[c] is None) else num_gt[c]
        rec = torch.zeros_like(tp) if denom == 0 else tp / denom          # meters.py:24-30 `div`: x / 0 := 0
        a = 0
        for t in thr:
            inds = torch.nonzero(rec >= t).squeeze()
            if inds.numel():
                a += prec[inds].max() / 11
        ap[c] = a
        max_rec[c] = rec[-1]
    return ap, max_rec


def synthetic_meter_stream(seed: int, num_cls: int = 600, batches: int = 12, per_batch: int = 4000, with_invalid: bool = False):
    """A sweep's worth of (scores, predictions, labels) appends with DISTINCT scores per class (the reference's unstable
    argsort leaves ties unspecified), a few classes never predicted, a few without any true positive, and num_gt
    >= the collected positives (test helper shared by the golden generator and the tests)."""
    g = torch.Generator().manual_seed(seed)
    stream = []
    perm = torch.randperm(batches * per_batch, generator=g).float()
    base = (perm + 0.5) / (batches * per_batch)                       # all distinct, in (0, 1), exact in fp32
    for b in range(batches):
        sc = base[b * per_batch: (b + 1) * per_batch].clone()
        pr = torch.randint(0, num_cls - 5, (per_batch,), generator=g).double()       # the last 5 classes stay empty
        lb = (torch.rand(per_batch, generator=g) < 0.2 * sc).float()                 # better scores -> more positives
        lb[pr < 3] = 0                                                               # classes 0..2: no true positive
        if with_invalid:                          # (the reference itself would fail on these: NaN.long() is no class index)
            pr[::97] = float("nan")
        stream.append((sc, pr, lb))
    pos = torch.zeros(num_cls)
    for sc, pr, lb in stream:
        ok = ~torch.isnan(pr)
        pos.index_add_(0, pr[ok].long(), lb[ok])
    num_gt = (pos + torch.randint(0, 4, (num_cls,), generator=g).float()).clamp(min=1).tolist()
    return stream, num_gt


def group_by_class(stream, num_cls: int):
    """What DetectionAPMeter.append accumulates (meters.py:585-604): per class, scores / labels in arrival order."""
    sc = [[] for _ in range(num_cls)]
    lb = [[] for _ in range(num_cls)]
    for scores, pred, labels in stream:
        ok = ~torch.isnan(pred.double())
        p = pred[ok].long()
        for c in p.unique().tolist():
            if 0 <= c < num_cls:
                sel = p == c
                sc[c].append(scores[ok][sel].double())
                lb[c].append(labels[ok][sel].double())
    cat = lambda xs: torch.cat(xs) if xs else torch.zeros(0, dtype=torch.float64)
    return [cat(x) for x in sc], [cat(x) for x in lb]


def test_hico_ref(outputs_per_batch, targets_per_batch, conversion: Optional[torch.Tensor], num_gt=None, num_cls: int = 600,
                  min_iou: float = 0.5):
    """The sweep of T:348-411 on CPU detections: per batch the per-image association above, the meter's per-class gathering
    (meters.py:585-604) and its 11-point AP (ap_11point).  -> ap (num_cls,) fp64."""
    scores = [[] for _ in range(num_cls)]
    labels = [[] for _ in range(num_cls)]
    for outputs, targets in zip(outputs_per_batch, targets_per_batch):
        if outputs is None or len(outputs) == 0:
            continue
        for output, target in zip(outputs, targets):
            s, inter, lab = associate_image(output, target, conversion, min_iou)
            for c in inter.unique().tolist():
                if c != c or c < 0 or c >= num_cls:          # NaN / invalid combinations are never gathered
                    continue
                sel = inter == c
                scores[int(c)].append(s[sel])
                labels[int(c)].append(lab[sel])
    cat = lambda xs: torch.cat(xs) if xs else torch.zeros(0)
    return ap_11point([cat(x) for x in scores], [cat(x) for x in labels], num_gt)[0]


def summarize_map_ref(ap: torch.Tensor, num_anno, unseen_idx=None) -> dict:
    """M:915-948."""
    num_anno = torch.as_tensor(num_anno)
    rare = torch.nonzero(num_anno < 10).squeeze(1)
    non_rare = torch.nonzero(num_anno >= 10).squeeze(1)
    out = {"full": float(ap.mean() * 100), "rare": float(ap[rare].mean() * 100), "non_rare": float(ap[non_rare].mean() * 100)}
    if unseen_idx is not None:
        ap_unseen = [v for i, v in enumerate(ap) if i in unseen_idx]
        ap_seen = [v for i, v in enumerate(ap) if i not in unseen_idx]
        out["unseen"] = float(torch.as_tensor(ap_unseen).mean() * 100)
        out["seen"] = float(torch.as_tensor(ap_seen).mean() * 100)
    return out
