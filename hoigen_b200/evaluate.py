"""Batched evaluation-side association: the consumer of the detections.

The reference's `CustomisedDLE.test_hico` (utils_tip_cache_and_union_finetune.py:348-411) copies every image's
detections to the host and loops in Python over images and HOI classes, calling `BoxPairAssociation`
(pocket/pocket/utils/association.py:51-125) for each.  `HOIAssociator` does the same for a whole batch with ONE
kernel launch on the packed detections the forward already produced (`DetectionList.packed`), and returns per image
exactly what the reference hands to `DetectionAPMeter.append(scores, interactions, labels)`.

    assoc = HOIAssociator(dataset.object_n_verb_to_interaction)         # once
    dets = upt.forward_from_proposals(...)                              # or upt(images)
    for scores, interactions, labels in assoc(dets, targets):           # targets: the reference's dicts
        meter.append(scores, interactions, labels)

No CPU fallback: CUDA tensors only (the compute is `hoigen_associate_pairs`).
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import torch

from . import _cabi

MAX_GT_PER_IMAGE = 1024


def recover_boxes(boxes: torch.Tensor, size: torch.Tensor) -> torch.Tensor:
    """UPT.recover_boxes (upt_tip_cache_model_free_finetune_distill3.py:1269-1274): normalised cxcywh -> xyxy pixels."""
    cx, cy, w, h = boxes.unbind(-1)
    xyxy = torch.stack([cx - 0.5 * w, cy - 0.5 * h, cx + 0.5 * w, cy + 0.5 * h], dim=-1)
    ih, iw = size
    return xyxy * torch.stack([iw, ih, iw, ih]).to(xyxy.dtype)


class HOIAssociator:
    """`conversion[objects, verbs]` + per-class `BoxPairAssociation(min_iou)` of the reference, batched on the GPU.

    object_n_verb_to_interaction: the dataset's 80 x num_verbs table (None / -1 = not an HOI class), or None when the
    detector's classes already are HOI ids (the reference's `else: interactions = verbs` branch, T:389-390)."""

    def __init__(self, object_n_verb_to_interaction: Optional[Sequence[Sequence[Optional[int]]]], min_iou: float = 0.5,
                 return_float_interactions: bool = True):
        self.min_iou = float(min_iou)
        self.return_float = return_float_interactions
        if object_n_verb_to_interaction is None:
            self._table_host, self.num_verbs = None, 1
        else:
            rows = [[-1 if (v is None or v < 0) else int(v) for v in row] for row in object_n_verb_to_interaction]
            if len(rows) != 80:
                raise ValueError(f"object_n_verb_to_interaction must have 80 rows (got {len(rows)})")
            self._table_host = torch.tensor(rows, dtype=torch.int32)
            self.num_verbs = self._table_host.shape[1]
        self._table_dev = {}

    def _table(self, dev):
        if self._table_host is None:
            return None
        t = self._table_dev.get(dev)
        if t is None:
            t = self._table_dev[dev] = self._table_host.to(dev).contiguous()
        return t

    @torch.no_grad()
    def __call__(self, detections, targets: Sequence[dict]) -> List[Tuple[torch.Tensor, torch.Tensor, torch.Tensor]]:
        """detections: the DetectionList of a forward (its `.packed` record is used as is) or a List[dict] in the
        reference's format; targets: per image {boxes_h, boxes_o (G,4) normalised cxcywh, hoi (G,), size (2,) = (h, w)}.
        Returns per image (scores, interactions, labels); interactions are float with NaN for invalid combinations when
        `return_float_interactions` (the reference's dtype), else int64 with -1."""
        packed = getattr(detections, "packed", None)
        if packed is None:
            packed = _pack(detections)
        dev = packed.scores.device
        if dev.type != "cuda":
            raise _cabi.HoigenError("HOIAssociator needs CUDA tensors (there is no CPU fallback)")
        _cabi.init(dev)
        B = packed.num_images
        if len(targets) != B:
            raise ValueError(f"{B} images of detections, {len(targets)} targets")
        # ground truth -> detection frame, CSR over images: ONE concatenation / upload / recover_boxes for the batch (the
        # same fp32 operations per box as the reference's per-image recover_boxes calls)
        counts = [int(t["hoi"].numel()) for t in targets]
        goff = [0]
        for c in counts:
            goff.append(goff[-1] + c)
        max_gt = max(counts) if counts else 0
        if max_gt > MAX_GT_PER_IMAGE:
            raise ValueError(f"at most {MAX_GT_PER_IMAGE} ground-truth pairs per image (got {max_gt})")
        if goff[-1]:
            raw = torch.cat([torch.cat([t["boxes_h"].reshape(-1, 4).float(), t["boxes_o"].reshape(-1, 4).float()], dim=1)
                             for t in targets]).to(dev, non_blocking=True)                     # (G, 8) cxcywh | cxcywh
            sizes_hw = torch.stack([t["size"].reshape(2).float() for t in targets]).to(dev, non_blocking=True)   # (B, 2) = (h, w)
            per_gt = sizes_hw.repeat_interleave(torch.tensor(counts, device=dev), dim=0)                 # (G, 2)
            scale = torch.stack([per_gt[:, 1], per_gt[:, 0], per_gt[:, 1], per_gt[:, 0]], dim=1)         # (w, h, w, h)

            def rec(b):
                cx, cy, w, h = b.unbind(-1)
                return (torch.stack([cx - 0.5 * w, cy - 0.5 * h, cx + 0.5 * w, cy + 0.5 * h], dim=-1) * scale).contiguous()
            gt_h, gt_o = rec(raw[:, :4]), rec(raw[:, 4:])
            gt_hoi = torch.cat([t["hoi"].reshape(-1).to(torch.int64) for t in targets]).to(dev, non_blocking=True).contiguous()
        else:
            gt_h = gt_o = torch.zeros(1, 4, device=dev)
            gt_hoi = torch.zeros(1, dtype=torch.int64, device=dev)
        offs = torch.tensor(packed.box_off + packed.triplet_off + goff, dtype=torch.int32)
        d_offs = torch.empty_like(offs, device=dev)
        _cabi.call("hoigen_set_words", d_offs.data_ptr(), offs.data_ptr(), offs.numel())
        d_box_off, d_trip_off, d_gt_off = d_offs[: B + 1], d_offs[B + 1: 2 * B + 2], d_offs[2 * B + 2:]
        mtot = int(packed.scores.numel())
        interactions = torch.empty(max(mtot, 1), dtype=torch.int64, device=dev)
        labels = torch.empty(max(mtot, 1), dtype=torch.float32, device=dev)
        table = self._table(dev)
        boxes = packed.boxes.float().contiguous()
        if mtot:
            _cabi.call("hoigen_associate_pairs", boxes.data_ptr(), d_box_off.data_ptr(), packed.pairing.data_ptr(),
                       packed.objects.data_ptr(), packed.labels.data_ptr(), packed.scores.data_ptr(), d_trip_off.data_ptr(),
                       table.data_ptr() if table is not None else None, self.num_verbs, gt_h.data_ptr(), gt_o.data_ptr(),
                       gt_hoi.data_ptr(), d_gt_off.data_ptr(), max_gt, B, self.min_iou, interactions.data_ptr(),
                       labels.data_ptr())
        inter = interactions[:mtot]
        if self.return_float:
            inter = torch.where(inter < 0, torch.full((), float("nan"), device=dev, dtype=torch.float64), inter.double())
        sizes = [b - a for a, b in zip(packed.triplet_off[:-1], packed.triplet_off[1:])]
        return list(zip(packed.scores.split(sizes), inter.split(sizes), labels[:mtot].split(sizes)))


def _pack(dets: Sequence[Optional[dict]]):
    """List[dict] in the reference's format -> PackedDetections (images without detections get zero-length segments)."""
    from .detector import PackedDetections
    dev = next(d["scores"].device for d in dets if d is not None)
    toff, boff = [0], [0]
    sc, lb, ob, pr, bx = [], [], [], [], []
    for d in dets:
        m = 0 if d is None else int(d["scores"].numel())
        n = 0 if d is None else int(d["boxes"].shape[0])
        toff.append(toff[-1] + m)
        boff.append(boff[-1] + n)
        if d is not None:
            sc.append(d["scores"].float()); lb.append(d["labels"].to(torch.int64)); ob.append(d["objects"].to(torch.int64))
            pr.append(d["pairing"].to(torch.int64).reshape(-1)); bx.append(d["boxes"].float().view(-1, 4))
    cat = lambda ts, dt, shape=(0,): torch.cat(ts) if ts else torch.zeros(shape, dtype=dt, device=dev)
    return PackedDetections(cat(sc, torch.float32), cat(lb, torch.int64), cat(ob, torch.int64), cat(pr, torch.int64),
                            cat(bx, torch.float32, (0, 4)), toff, boff, (0, 0))


class DetectionAPMeter:
    """GPU counterpart of pocket.utils.DetectionAPMeter(num_cls, num_gt, algorithm='11P', precision=64)
    (pocket/pocket/utils/meters.py:414-639) with the same append / eval / reset surface.  Results stay on the device
    between `append`s (no `.tolist()` per class per image); `eval` orders the sweep once by (class, score) and computes
    every class's precision / recall curve and 11-point AP in one kernel (`hoigen_ap_11point`).

    Differences from the reference, both deliberate: detections of one class with EQUAL scores keep their arrival order
    (the reference's `argsort` is unstable, i.e. its order among ties is unspecified), and invalid class ids (negative or
    NaN `prediction`s) are dropped as the reference's per-class gathering effectively does."""

    def __init__(self, num_cls: int, num_gt=None, algorithm: str = "11P", nproc: int = 1, precision: int = 64):
        if algorithm != "11P":
            raise NotImplementedError("hoigen_b200 implements the '11P' algorithm the reference's eval uses (T:363-367)")
        if precision != 64:
            raise NotImplementedError("fp64 only (the reference's default precision)")
        if num_gt is not None and len(num_gt) != num_cls:
            raise AssertionError("Provided ground truth instances do not have the same number of classes as specified")
        self.num_cls = num_cls
        self.num_gt = None if num_gt is None else [(-1.0 if g is None else float(g)) for g in num_gt]
        self.algorithm = algorithm
        self.reset()

    def reset(self, keep_old: bool = False) -> None:
        if not keep_old:
            self._scores, self._classes, self._labels = [], [], []

    def append(self, output: torch.Tensor, prediction: torch.Tensor, labels: torch.Tensor) -> None:
        """output (N,) scores, prediction (N,) class ids (float with NaN or int), labels (N,) 0/1 — meters.py:585-604."""
        if not (isinstance(output, torch.Tensor) and isinstance(prediction, torch.Tensor) and isinstance(labels, torch.Tensor)):
            raise TypeError("Arguments should be torch.Tensor")
        if output.device.type != "cuda":
            raise _cabi.HoigenError("DetectionAPMeter needs CUDA tensors (there is no CPU fallback)")
        pred = prediction
        if pred.is_floating_point():
            pred = torch.where(torch.isnan(pred), torch.full_like(pred, -1.0), pred)
        self._scores.append(output.detach().float().reshape(-1))
        self._classes.append(pred.detach().long().reshape(-1))
        self._labels.append(labels.detach().float().reshape(-1))

    @torch.no_grad()
    def eval(self) -> torch.Tensor:
        """-> ap (num_cls,) fp64 on the device; also sets .ap and .max_rec like the reference (meters.py:621-639)."""
        if not self._scores:
            raise RuntimeError("nothing has been appended")
        dev = self._scores[0].device
        _cabi.init(dev)
        scores, classes, labels = torch.cat(self._scores), torch.cat(self._classes), torch.cat(self._labels)
        keep = (classes >= 0) & (classes < self.num_cls)
        scores, classes, labels = scores[keep], classes[keep], labels[keep]
        # (class ascending, score descending), ties in arrival order: two stable sorts
        o1 = torch.argsort(scores, descending=True, stable=True)
        o2 = torch.argsort(classes[o1], stable=True)
        order = o1[o2]
        lab_sorted = labels[order].contiguous()
        counts = torch.bincount(classes, minlength=self.num_cls)
        class_off = torch.zeros(self.num_cls + 1, dtype=torch.int64, device=dev)
        class_off[1:] = counts.cumsum(0)
        if self.num_gt is not None:
            tp = torch.zeros(self.num_cls, dtype=torch.float64, device=dev).index_add_(0, classes, labels.double())
            ngt = torch.tensor(self.num_gt, dtype=torch.float64, device=dev)
            bad = torch.nonzero((ngt >= 0) & (tp > ngt)).flatten()
            if bad.numel():
                raise AssertionError(f"Class {int(bad[0])}: Number of true positives larger than that of ground truth")
        else:
            ngt = torch.full((self.num_cls,), -1.0, dtype=torch.float64, device=dev)
        thr = torch.linspace(0, 1, 11, dtype=torch.float64).to(dev)       # the reference's own threshold values
        ap = torch.empty(self.num_cls, dtype=torch.float64, device=dev)
        max_rec = torch.empty(self.num_cls, dtype=torch.float64, device=dev)
        _cabi.call("hoigen_ap_11point", lab_sorted.data_ptr() if lab_sorted.numel() else None, class_off.data_ptr(),
                   ngt.data_ptr(), thr.data_ptr(), self.num_cls, ap.data_ptr(), max_rec.data_ptr())
        self.ap, self.max_rec = ap, max_rec
        return ap


@torch.no_grad()
def test_hico(net, dataloader, object_n_verb_to_interaction=None, num_gt=None, *, tgt_num_classes: int = 600,
              min_iou: float = 0.5, to_device=None, launch_ahead: bool = True) -> torch.Tensor:
    """The reference's evaluation sweep `CustomisedDLE.test_hico` (utils_tip_cache_and_union_finetune.py:348-411) around
    the accelerated forward: for every batch `(inputs, ..., targets)` of the loader run `net(inputs, targets)`, skip
    batches without detections (T:371-373), convert (object, verb) -> HOI id, associate with the ground truth and feed
    `DetectionAPMeter(600, num_gt, '11P')`; returns `meter.eval()` (per-class AP, fp64, on the device).

    Differences from the reference, none of which changes a number: any batch size (the reference is fixed at 1), the
    detections never leave the GPU (association and AP are the batched kernels above), and when `net` offers
    `launch(inputs, targets)` / `finish(handle)` (hoigen_b200.detector.UPT: `launch_from_proposals` / `finish` behind a thin
    adapter) the next batch is enqueued before the previous one is post-processed.

    net: callable `(inputs, targets) -> List[dict] | None` like the reference's detector (U:1543).
    object_n_verb_to_interaction: `dataset.object_n_verb_to_interaction` for 117-verb models (T:387-388); None when the
    model's classes already are HOI ids (T:389-390).  num_gt: `dataset.anno_interaction` (HICO-DET) or None (T:361)."""
    assoc = HOIAssociator(object_n_verb_to_interaction, min_iou=min_iou)
    meter = DetectionAPMeter(tgt_num_classes, num_gt=num_gt, algorithm="11P")
    pipelined = launch_ahead and hasattr(net, "launch") and hasattr(net, "finish")

    def consume(outputs, targets):
        if outputs is None or len(outputs) == 0:      # T:371-373
            return
        for scores, interactions, labels in assoc(outputs, targets):
            meter.append(scores, interactions, labels)

    pending = None
    for batch in dataloader:
        inputs, targets = batch[0], batch[-1]
        if to_device is not None:
            inputs = to_device(inputs)
        if pipelined:
            handle = net.launch(inputs, targets)
            if pending is not None:
                consume(net.finish(pending[0]), pending[1])
            pending = (handle, targets)
        else:
            consume(net(inputs, targets), targets)
    if pending is not None:
        consume(net.finish(pending[0]), pending[1])
    return meter.eval()


def summarize_map(ap: torch.Tensor, num_anno=None, unseen_hoi_idx: Optional[Sequence[int]] = None) -> dict:
    """The numbers main_tip_finetune.py:915-948 prints: full mAP, rare (< 10 training annotations) / non-rare, and for the
    zero-shot settings seen / unseen over `hico_unseen_index[zs_type]`.  Values in percent, like the reference's log."""
    ap = ap.detach().double().cpu()
    out = {"full": float(ap.mean() * 100)}
    if num_anno is not None:
        na = torch.as_tensor(num_anno)
        rare, non_rare = torch.nonzero(na < 10).squeeze(1), torch.nonzero(na >= 10).squeeze(1)
        out["rare"] = float(ap[rare].mean() * 100)
        out["non_rare"] = float(ap[non_rare].mean() * 100)
    if unseen_hoi_idx is not None:
        unseen = sorted(set(int(i) for i in unseen_hoi_idx))
        seen = [i for i in range(ap.numel()) if i not in set(unseen)]
        out["unseen"] = float(ap[torch.tensor(unseen)].mean() * 100)
        out["seen"] = float(ap[torch.tensor(seen)].mean() * 100)
    return out
