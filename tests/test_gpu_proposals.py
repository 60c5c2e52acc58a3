"""f3 on the GPU: hoigen_prepare_proposals (through UPT.prepare_region_proposals_batched) against the torchvision-free
restatement, the reference's committed outputs and the per-image torch form.  Bar: selection, order and values BIT-EXACT."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _build(dev, min_instances=3, max_instances=15, N=128):
    from hoigen_b200 import synthetic as S
    from hoigen_b200.detector import UPT
    enc = S.make_encoder_state(0)
    head = S.make_head_state(117, N, seed=2, max_instances=max_instances)
    m = UPT.from_state(enc, head).to(dev)
    m.min_instances, m.max_instances = min_instances, max_instances
    return m, enc, head


def _to(results, dev):
    return [{k: v.to(dev) for k, v in r.items()} for r in results]


def _split(batched):
    boxes, scores, labels, n_list, nh_list = batched
    out, s = [], 0
    for n, nh in zip(n_list, nh_list):
        out.append(dict(boxes=boxes[s:s + n].cpu(), scores=scores[s:s + n].cpu(), labels=labels[s:s + n].cpu(), n_human=nh))
        s += n
    assert s == boxes.shape[0]
    return out


def test_kernel_matches_reference_golden(cuda_device):
    """tests/golden/proposals.npz = the UNMODIFIED reference's prepare_region_proposals (ragged candidate counts:
    one launch per image here)."""
    m, _, _ = _build(cuda_device)
    gold = np.load("tests/golden/proposals.npz")
    for b in range(4):
        res = [dict(scores=torch.from_numpy(gold[f"in_scores_{b}"]), labels=torch.from_numpy(gold[f"in_labels_{b}"]),
                    boxes=torch.from_numpy(gold[f"in_boxes_{b}"]))]
        got = _split(m.prepare_region_proposals_batched(_to(res, cuda_device)))[0]
        assert np.array_equal(got["boxes"].numpy(), gold[f"boxes_{b}"]), b
        assert np.array_equal(got["scores"].numpy(), gold[f"scores_{b}"]), b
        assert np.array_equal(got["labels"].numpy(), gold[f"labels_{b}"]), b
        assert got["n_human"] == int((gold[f"labels_{b}"] == 0).sum())


@pytest.mark.parametrize("seed,B,q,ties,lim", [(31, 64, 100, 0, (3, 15)), (32, 16, 100, 4, (3, 15)), (33, 8, 256, 0, (0, 16)),
                                                (34, 5, 1, 0, (3, 15)), (35, 12, 37, 9, (1, 4))])
def test_kernel_matches_oracle(cuda_device, seed, B, q, ties, lim):
    """Seeded DETR-like candidates (clusters of overlapping boxes, images without humans, all-below / all-above the score
    threshold, tied scores, 1 and 256 candidates) against the restatement, which keeps the stable NMS order for ties."""
    from oracle import hoi_forward_ref as O
    m, _, _ = _build(cuda_device, *lim)
    results = O.synthetic_detr_results(B, seed, q, ties)
    ref = O.prepare_region_proposals_ref(results, 0, 0.2, *lim)
    got = _split(m.prepare_region_proposals_batched(_to(results, cuda_device)))
    assert len(got) == B
    for b, (r, g) in enumerate(zip(ref, got)):
        assert torch.equal(r["boxes"], g["boxes"]) and torch.equal(r["scores"], g["scores"]), b
        assert torch.equal(r["labels"], g["labels"]) and r["n_human"] == g["n_human"], b
    if q >= 37:
        assert any(len(r["boxes"]) == 2 * lim[1] for r in ref) or lim[1] >= 15
        assert any(r["n_human"] == 0 for r in ref)


def test_kernel_matches_torch_form_on_device(cuda_device):
    """Same batch through the per-image torch form on the GPU (torchvision's CUDA nms) — no tied scores."""
    from oracle import hoi_forward_ref as O
    m, _, _ = _build(cuda_device)
    results = _to(O.synthetic_detr_results(32, 41, 100), cuda_device)
    ref = m.prepare_region_proposals(results)
    got = _split(m.prepare_region_proposals_batched(results))
    for b, (r, g) in enumerate(zip(ref, got)):
        assert torch.equal(r["boxes"].cpu(), g["boxes"]) and torch.equal(r["scores"].cpu(), g["scores"]), b
        assert torch.equal(r["labels"].cpu(), g["labels"]) and int(r["n_human"]) == g["n_human"], b


def test_ragged_or_cpu_batches_fall_back_to_none(cuda_device):
    from oracle import hoi_forward_ref as O
    m, _, _ = _build(cuda_device)
    a = O.synthetic_detr_results(2, 51, 100)
    assert m.prepare_region_proposals_batched(a) is None                        # CPU tensors
    ragged = _to(a, cuda_device)
    ragged[1] = {k: v[:50] for k, v in ragged[1].items()}
    assert m.prepare_region_proposals_batched(ragged) is None                   # different candidate counts
    assert m.prepare_region_proposals_batched(_to(O.synthetic_detr_results(1, 52, 300), cuda_device)) is None
    assert m.prepare_region_proposals_batched([]) is None


def test_upt_forward_uses_batched_proposals(cuda_device):
    """UPT.forward with a stub DETR emitting a uniform (B, 100) candidate batch takes the one-kernel proposal stage; the
    detections equal forward_from_proposals over the per-image torch form, every field bit for bit."""
    from hoigen_b200 import synthetic as S
    from oracle import hoi_forward_ref as O
    m, enc, head = _build(cuda_device, N=256)
    B = 6
    results = _to(O.synthetic_detr_results(B, 61, 100), cuda_device)

    class Stub(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.query_embed = torch.nn.Embedding(1, 1)
            self.class_embed = self.bbox_embed = self.input_proj = torch.nn.Identity()

        def backbone(self, nested):
            from hoigen_b200.detector import _NestedTensor
            return [_NestedTensor(nested.tensors[:, :1, :1, :1], None)], [None]

        def transformer(self, src, mask, query, pos):
            return torch.zeros(1, src.shape[0], 1, 4, device=src.device), None

    class PP(torch.nn.Module):
        def forward(self, outputs, sizes):
            return results

    m.detector, m.postprocessor = Stub().to(cuda_device), PP()
    imgs = S.make_images(B, seed=5).to(cuda_device)
    dino = S.make_dino_features(B).to(cuda_device)
    m.dino_model = lambda x: dino
    calls = []
    orig = m.prepare_region_proposals_batched
    m.prepare_region_proposals_batched = lambda r: calls.append(1) or orig(r)
    dets = m([(torch.zeros(3, 40, 50, device=cuda_device), imgs[b]) for b in range(B)])
    assert calls == [1]
    ref = m.forward_from_proposals(imgs, m.prepare_region_proposals(results), dino)
    assert len(dets) == len(ref) == B
    for b in range(B):
        for k in ("boxes", "pairing", "labels", "objects", "scores", "size"):
            assert torch.equal(dets[b][k], ref[b][k]), (b, k)
    assert sum(int(d["scores"].numel()) for d in dets) > 1000
