"""f3 (proposal stage, the part that is not the DETR network): the torchvision-free restatement
oracle/hoi_forward_ref.py::prepare_region_proposals_ref against (1) what the UNMODIFIED reference produced
(tests/golden/proposals.npz, oracle/make_golden.py::proposals_case) and (2) the installed torchvision batched_nms on seeded
DETR-like candidates.  CPU; the CUDA kernel is checked in test_gpu_proposals.py."""
import numpy as np
import pytest
import torch

from oracle import hoi_forward_ref as O


def _same(a, b):
    return torch.equal(a["boxes"], b["boxes"]) and torch.equal(a["scores"], b["scores"]) and torch.equal(a["labels"], b["labels"])


def test_restatement_matches_reference_golden():
    gold = np.load("tests/golden/proposals.npz")
    results = [dict(scores=torch.from_numpy(gold[f"in_scores_{b}"]), labels=torch.from_numpy(gold[f"in_labels_{b}"]),
                    boxes=torch.from_numpy(gold[f"in_boxes_{b}"])) for b in range(4)]
    rp = O.prepare_region_proposals_ref(results, 0, 0.2, 3, 15)
    for b in range(4):
        assert np.array_equal(rp[b]["boxes"].numpy(), gold[f"boxes_{b}"]), b
        assert np.array_equal(rp[b]["scores"].numpy(), gold[f"scores_{b}"]), b
        assert np.array_equal(rp[b]["labels"].numpy(), gold[f"labels_{b}"]), b
    assert [r["n_human"] for r in rp] == [int((gold[f"labels_{b}"] == 0).sum()) for b in range(4)] == [1, 15, 5, 0]


@pytest.mark.parametrize("seed,q,ties", [(5, 100, 0), (6, 100, 7), (7, 37, 0), (8, 256, 0)])
def test_nms_restatement_matches_torchvision(seed, q, ties):
    from torchvision.ops.boxes import batched_nms
    for r in O.synthetic_detr_results(4, seed, q, ties):
        ref = batched_nms(r["boxes"], r["scores"], r["labels"], 0.5)
        mine = O.batched_nms_ref(r["boxes"], r["scores"], r["labels"], 0.5)
        assert torch.equal(ref, mine)
        assert 0 < len(mine) < q                                   # NMS suppressed something, kept something


@pytest.mark.parametrize("seed,q,ties,lim", [(15, 100, 0, (3, 15)), (16, 100, 0, (2, 6)), (17, 64, 0, (1, 4)), (18, 100, 0, (0, 16))])
def test_selection_restatement_matches_torchvision_form(seed, q, ties, lim):
    """Every branch of U:1374-1395 (fewer than min, more than max, in between; no human at all) is a prefix of the
    descending-score order that NMS leaves — the closed form the kernel uses.  (No tied scores here: the reference's
    `argsort(descending=True)` is not a stable sort, so the order among EQUAL scores in the min / max branches is
    whatever torch's sort does; the restatement and the kernel keep the stable NMS order.)"""
    results = O.synthetic_detr_results(8, seed, q, ties)
    ref = O.prepare_region_proposals(results, 0, 0.2, *lim)
    mine = O.prepare_region_proposals_ref(results, 0, 0.2, *lim)
    for b, (a, m) in enumerate(zip(ref, mine)):
        assert _same(a, m), b
        assert m["n_human"] == int((m["labels"] == 0).sum())
    assert mine[1]["n_human"] == 0 and len(mine[1]["boxes"]) > 0
