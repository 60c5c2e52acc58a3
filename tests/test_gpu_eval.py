"""f1 on the GPU: hoigen_associate_pairs (through hoigen_b200.evaluate.HOIAssociator) against the oracle restatement
and the reference's committed outputs.  Bar: interactions and labels BIT-EXACT."""
import json
from pathlib import Path

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

CASES = {"assoc_b6": dict(batch=6, seed=11, n_h=4, n_o=4), "assoc_b3_dense": dict(batch=3, seed=12, n_h=6, n_o=2)}


def _tables():
    from hoigen_b200 import synthetic as S
    return json.load(open(Path(S.__file__).parent / "data" / "object_tables.json"))["hico_object_n_verb_to_interaction"]


def _to(dets, dev):
    return [{k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in d.items()} for d in dets]


@pytest.mark.parametrize("name", list(CASES))
def test_association_matches_reference_golden(cuda_device, name):
    from hoigen_b200.evaluate import HOIAssociator
    from oracle import eval_ref as E
    c = CASES[name]
    gold = np.load(f"tests/golden/{name}.npz")
    onv = _tables()
    conv = E.conversion_table(onv)
    dets = E.synthetic_detections(c["batch"], c["seed"], c["n_h"], c["n_o"])
    tgts = E.make_targets(dets, conv, seed=c["seed"] + 1000)
    res = HOIAssociator(onv)(_to(dets, cuda_device), tgts)
    assert len(res) == c["batch"]
    for b, (scores, inter, labels) in enumerate(res):
        assert torch.equal(scores.cpu(), dets[b]["scores"])
        assert np.array_equal(np.nan_to_num(inter.cpu().numpy(), nan=-1.0), np.nan_to_num(gold[f"interactions_{b}"], nan=-1.0))
        assert np.array_equal(labels.cpu().numpy(), gold[f"labels_{b}"]), b


def test_association_matches_oracle_random_and_edges(cuda_device):
    """Fresh seeds, B = 16; one image without ground truth, one without detections, heavy score ties; also the
    `interactions = verbs` branch (no conversion table)."""
    from hoigen_b200.evaluate import HOIAssociator
    from oracle import eval_ref as E
    onv = _tables()
    conv = E.conversion_table(onv)
    dets = E.synthetic_detections(16, 77, 5, 3, tie_every=3)
    tgts = E.make_targets(dets, conv, seed=78, per_image=10)
    tgts[3] = dict(boxes_h=torch.zeros(0, 4), boxes_o=torch.zeros(0, 4), hoi=torch.zeros(0, dtype=torch.int64), size=tgts[3]["size"])
    empty = {k: (v[:0] if k in ("scores", "labels", "objects") else v) for k, v in dets[5].items()}
    empty["pairing"] = dets[5]["pairing"][:, :0]
    dets[5] = empty
    ref = E.associate_batch(dets, tgts, conv)
    got = HOIAssociator(onv)(_to(dets, cuda_device), tgts)
    for b, ((rs, ri, rl), (gs, gi, gl)) in enumerate(zip(ref, got)):
        assert torch.equal(ri.nan_to_num(-1), gi.cpu().nan_to_num(-1)), b
        assert torch.equal(rl, gl.cpu()), b
    assert sum(int(r[2].sum()) for r in ref) > 20
    # classes already are HOI ids: interactions = verbs (T:389-390); ground truth ids drawn from the verbs
    tg2 = [dict(t, hoi=(d["labels"][:t["hoi"].numel()] if d["labels"].numel() >= t["hoi"].numel() else t["hoi"]))
           for t, d in zip(tgts, dets)]
    ref2 = E.associate_batch(dets, tg2, None)
    got2 = HOIAssociator(None)(_to(dets, cuda_device), tg2)
    for (rs, ri, rl), (gs, gi, gl) in zip(ref2, got2):
        assert torch.equal(ri.double(), gi.cpu().double()) and torch.equal(rl, gl.cpu())


def test_association_on_forward_output(cuda_device):
    """End to end: the detections of a real forward (DetectionList.packed, no re-packing) associated with ground truth
    derived from them; equals the oracle run on the same detections copied to the host."""
    from hoigen_b200 import synthetic as S
    from hoigen_b200.detector import UPT
    from hoigen_b200.evaluate import HOIAssociator
    from oracle import eval_ref as E
    onv = _tables()
    conv = E.conversion_table(onv)
    m = UPT.from_state(S.make_encoder_state(0), S.make_head_state(117, 256, seed=2)).to(cuda_device)
    B = 4
    props = [{k: v.to(cuda_device) for k, v in p.items()} for p in S.make_region_props(B, 4, 4, seed=90)]
    dets = m.forward_from_proposals(S.make_images(B, seed=91).to(cuda_device), props, S.make_dino_features(B, seed=92).to(cuda_device))
    host = [{k: (v.cpu() if torch.is_tensor(v) else v) for k, v in d.items()} for d in dets]
    tgts = E.make_targets(host, conv, seed=93)
    ref = E.associate_batch(host, tgts, conv)
    got = HOIAssociator(onv)(dets, tgts)
    for (rs, ri, rl), (gs, gi, gl) in zip(ref, got):
        assert torch.equal(ri.nan_to_num(-1), gi.cpu().nan_to_num(-1)) and torch.equal(rl, gl.cpu())
    assert sum(int(r[2].sum()) for r in ref) > 0


@pytest.mark.parametrize("name,seed,with_gt", [("ap11_with_num_gt", 21, True), ("ap11_no_num_gt", 22, False)])
def test_ap_meter_matches_reference_golden(cuda_device, name, seed, with_gt):
    """hoigen_b200.evaluate.DetectionAPMeter (hoigen_ap_11point) against the committed outputs of the unmodified
    pocket.utils.DetectionAPMeter: fp64 AP and max recall of all 600 classes bit for bit."""
    from hoigen_b200.evaluate import DetectionAPMeter
    from oracle import eval_ref as E
    gold = np.load(f"tests/golden/{name}.npz")
    stream, num_gt = E.synthetic_meter_stream(seed)
    meter = DetectionAPMeter(600, num_gt=num_gt if with_gt else None, algorithm="11P")
    for sc, pr, lb in stream:
        meter.append(sc.to(cuda_device), pr.to(cuda_device), lb.to(cuda_device))
    ap = meter.eval()
    assert ap.dtype == torch.float64 and ap.shape == (600,)
    assert np.array_equal(ap.cpu().numpy(), gold["ap"])
    assert np.array_equal(meter.max_rec.cpu().numpy(), gold["max_rec"])


def test_ap_meter_edges(cuda_device):
    """Invalid class ids are dropped, ties keep arrival order (the oracle's stable order), a class larger than one
    scan chunk, too many true positives are refused like the reference does."""
    from hoigen_b200.evaluate import DetectionAPMeter
    from oracle import eval_ref as E
    stream, num_gt = E.synthetic_meter_stream(31, num_cls=40, batches=5, per_batch=3000, with_invalid=True)
    stream = [(torch.round(sc * 50) / 50, pr, lb) for sc, pr, lb in stream]          # heavy score ties
    sc, lb = E.group_by_class(stream, 40)
    ref_ap, ref_rec = E.ap_11point(sc, lb, num_gt)
    meter = DetectionAPMeter(40, num_gt=num_gt)
    for s, p, l in stream:
        meter.append(s.to(cuda_device), p.to(cuda_device), l.to(cuda_device))
    ap = meter.eval()
    assert torch.equal(ap.cpu(), ref_ap) and torch.equal(meter.max_rec.cpu(), ref_rec)
    bad = DetectionAPMeter(2, num_gt=[1, 5])
    bad.append(torch.tensor([0.9, 0.8, 0.7], device=cuda_device), torch.tensor([0, 0, 1], device=cuda_device),
               torch.tensor([1.0, 1.0, 0.0], device=cuda_device))
    with pytest.raises(AssertionError):
        bad.eval()


class _LoaderNet:
    """What main_tip_finetune.py hands to CustomisedDLE.test_hico, reduced to the accelerated path: `net(inputs, targets)`
    like the reference's detector, plus launch / finish so that evaluate.test_hico can keep a batch in flight."""

    def __init__(self, model, dev):
        self.model, self.dev = model, dev

    def launch(self, inputs, targets):
        imgs, props, dino = inputs
        return self.model.launch_from_proposals(imgs.to(self.dev), [{k: (v.to(self.dev) if torch.is_tensor(v) else v) for k, v in p.items()}
                                                                    for p in props], dino.to(self.dev))

    def finish(self, handle):
        return self.model.finish(handle)

    def __call__(self, inputs, targets):
        return self.finish(self.launch(inputs, targets))


def test_sweep_driver_map_matches_reference_loop(cuda_device):
    """f1 end to end (T:348-411 + M:915-948): a seeded synthetic 'dataset' of 6 batches x 4 images through
    evaluate.test_hico (forward -> HOIAssociator -> DetectionAPMeter -> per-class AP -> full / rare / non-rare / seen / unseen
    mAP) against the oracle's restatement of the reference's loop.  (a) On the SAME detections the 600 APs are bit-identical
    (fp64) whether batches are pipelined or not; (b) against the loop fed with the ORACLE forward's detections (fp32 CPU
    reference arithmetic) the mAPs agree to a small fraction of a point (the bf16 logits move a few near-tied scores)."""
    from hoigen_b200 import synthetic as S
    from hoigen_b200.detector import UPT
    from hoigen_b200.evaluate import summarize_map, test_hico
    from oracle import eval_ref as E
    from oracle import hoi_forward_ref as O
    onv = _tables()
    conv = E.conversion_table(onv)
    enc, head = S.make_encoder_state(0), S.make_head_state(117, 256, seed=2)
    model = UPT.from_state(enc, head).to(cuda_device)
    net = _LoaderNet(model, cuda_device)
    batches = []
    for i in range(6):
        B = 4
        props = S.make_region_props(B, 4, 4, ragged=(i % 2 == 1), seed=900 + 10 * i)
        if i == 3:
            for p in props:                       # a batch without any human: the forward returns None (T:371-373)
                p["labels"] = torch.full_like(p["labels"], 7)
        batches.append((S.make_images(B, seed=910 + i), props, S.make_dino_features(B, seed=920 + i)))
    # ground truth derived from a first pass over the data (jittered copies of detected pairs + misses), seeded
    dets_cpu, oracle_dets, targets = [], [], []
    for i, inp in enumerate(batches):
        out = net(inp, None)
        o_out = O.hoi_forward(inp[0], inp[1], inp[2], enc, head) if out is not None else None
        if out is None:
            dets_cpu.append(None); oracle_dets.append(None)
            targets.append([dict(boxes_h=torch.zeros(0, 4), boxes_o=torch.zeros(0, 4), hoi=torch.zeros(0, dtype=torch.int64),
                                 size=torch.tensor([224.0, 224.0])) for _ in range(4)])
            continue
        cpu = [{k: v.cpu() for k, v in d.items()} for d in out]
        dets_cpu.append(cpu); oracle_dets.append(o_out)
        targets.append(E.make_targets(cpu, conv, seed=930 + i, per_image=8))
    loader = [(inp, tg) for inp, tg in zip(batches, targets)]
    g = torch.Generator().manual_seed(5)
    num_anno = torch.randint(1, 40, (600,), generator=g)
    num_gt = [float(v) for v in (torch.randint(200, 400, (600,), generator=g)).tolist()]
    uc0 = S.load_object_tables()["hico_unseen_uc0"]
    ap_pipe = test_hico(net, loader, onv, num_gt=num_gt)
    ap_seq = test_hico(net, loader, onv, num_gt=num_gt, launch_ahead=False)
    assert torch.equal(ap_pipe, ap_seq)
    ref_same = E.test_hico_ref(dets_cpu, targets, conv, num_gt=num_gt)
    assert torch.equal(ap_pipe.cpu(), ref_same), (ap_pipe.cpu() - ref_same).abs().max()
    ours = summarize_map(ap_pipe, num_anno, uc0)
    ref = E.summarize_map_ref(ref_same, num_anno, uc0)
    assert ours == pytest.approx(ref, abs=1e-9) and ours["full"] > 0.05
    ref_oracle = E.summarize_map_ref(E.test_hico_ref(oracle_dets, targets, conv, num_gt=num_gt), num_anno, uc0)
    print("mAP (ours | reference loop on oracle-forward detections):",
          {k: (round(ours[k], 3), round(ref_oracle[k], 3)) for k in ours})
    for k in ours:
        assert abs(ours[k] - ref_oracle[k]) <= 0.25, (k, ours[k], ref_oracle[k])
