"""f1 (batched eval association): the oracle restatement against the reference's committed outputs and its own
known-answer test (CPU; the CUDA path is checked in test_gpu_eval.py)."""
import json
from pathlib import Path

import numpy as np
import pytest
import torch

from oracle import eval_ref as E

CASES = {"assoc_b6": dict(batch=6, seed=11, n_h=4, n_o=4), "assoc_b3_dense": dict(batch=3, seed=12, n_h=6, n_o=2)}


def conversion():
    from hoigen_b200 import synthetic as S
    tables = json.load(open(Path(S.__file__).parent / "data" / "object_tables.json"))
    return tables["hico_object_n_verb_to_interaction"]


@pytest.mark.parametrize("name", list(CASES))
def test_oracle_matches_reference_golden(name):
    """tests/golden/assoc_*.npz hold what the UNMODIFIED reference loop + pocket.utils.BoxPairAssociation produced
    (oracle/make_golden_eval.py) for these seeded inputs."""
    c = CASES[name]
    gold = np.load(f"tests/golden/{name}.npz")
    conv = E.conversion_table(conversion())
    dets = E.synthetic_detections(c["batch"], c["seed"], c["n_h"], c["n_o"])
    tgts = E.make_targets(dets, conv, seed=c["seed"] + 1000)
    res = E.associate_batch(dets, tgts, conv)
    tp = 0
    for b, (scores, inter, labels) in enumerate(res):
        assert np.array_equal(np.nan_to_num(inter.numpy(), nan=-1.0), np.nan_to_num(gold[f"interactions_{b}"], nan=-1.0))
        assert np.array_equal(labels.numpy(), gold[f"labels_{b}"])
        tp += int(labels.sum())
    assert tp > 0


def test_known_answer_duplicates():
    """pocket/test/association.py::test_duplicates: one ground truth, two overlapping detections — with scores the
    higher-scoring one is the true positive, without scores the higher-IoU one."""
    gt = torch.tensor([[30., 30., 60., 60.]])
    det = torch.tensor([[28.8, 31.2, 59.1, 58.4], [26.9, 29.2, 63.5, 66.4]])
    iou = E.box_iou(gt, det)
    assert E.box_association(iou, torch.tensor([0.8, 0.9]), 0.5).tolist() == [0., 1.]
    assert E.box_association(iou, None, 0.5).tolist() == [1., 0.]


def test_box_iou_matches_torchvision_bitwise():
    from torchvision.ops import box_iou
    g = torch.Generator().manual_seed(3)
    a = torch.rand(40, 4, generator=g) * 100
    b = torch.rand(55, 4, generator=g) * 100
    a[:, 2:] += a[:, :2]
    b[:, 2:] += b[:, :2]
    assert torch.equal(E.box_iou(a, b), box_iou(a, b))


def test_recover_boxes():
    b = torch.tensor([[0.5, 0.25, 0.2, 0.1]])
    out = E.recover_boxes(b, torch.tensor([200.0, 400.0]))
    assert torch.allclose(out, torch.tensor([[160.0, 40.0, 240.0, 60.0]]))


@pytest.mark.parametrize("name,seed,with_gt", [("ap11_with_num_gt", 21, True), ("ap11_no_num_gt", 22, False)])
def test_ap_oracle_matches_reference_golden(name, seed, with_gt):
    """tests/golden/ap11_*.npz: per-class AP / max recall of the UNMODIFIED pocket.utils.DetectionAPMeter('11P', fp64) on a
    seeded 48 000-detection sweep over 600 classes (empty classes, classes without true positives, num_gt given or
    not) — the restatement reproduces them bit for bit."""
    gold = np.load(f"tests/golden/{name}.npz")
    stream, num_gt = E.synthetic_meter_stream(seed)
    sc, lb = E.group_by_class(stream, 600)
    ap, rec = E.ap_11point(sc, lb, num_gt if with_gt else None)
    assert np.array_equal(ap.numpy(), gold["ap"])
    assert np.array_equal(rec.numpy(), gold["max_rec"])
    assert 0.1 < ap.mean().item() < 0.5
