"""Generate tests/golden/*.npz by running the UNMODIFIED reference (builder container only) and pin the oracle.

    python oracle/make_golden.py            # writes tests/golden/, prints oracle-vs-reference deviations

For every case the inputs are re-created from seeds (hoigen_b200/synthetic.py) — only OUTPUTS of the reference are
stored.  The same script asserts that oracle/hoi_forward_ref.py reproduces the reference stage by stage (fp32,
different summation order => small tolerances; indices bit-exact) and records the observed deviations in
tests/golden/PINNING.json, which DESIGN.md quotes.
"""
from __future__ import annotations

import json
import sys
import time
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

from hoigen_b200 import synthetic as S  # noqa: E402
from oracle import hoi_forward_ref as O  # noqa: E402
from oracle import ref_harness as RH  # noqa: E402

GOLD = ROOT / "tests" / "golden"

from oracle.golden_cases import CASES, head_for, props_for  # noqa: E402


def make_props(case):
    return props_for(case)


def run_case(name, case, refs):
    key = (case["num_classes"], case["dataset"], tuple(sorted(case.get("args", {}).items())))
    if key not in refs:
        t0 = time.time()
        over = dict(case.get("args", {}))
        refs[key] = RH.build_reference_upt(case["num_classes"], case["dataset"], **over)
        print(f"[{name}] built reference UPT in {time.time()-t0:.1f}s")
    upt, pp = refs[key]
    over = case.get("args", {})
    enc = S.make_encoder_state(0)
    head = head_for(case)
    RH.load_synthetic_state(upt, enc, head)
    imgs = S.make_images(case["B"], seed=1)
    props = make_props(case)
    dino = S.make_dino_features(case["B"])

    # ---- reference, whole forward (U:1543) ----
    ref_dets = RH.run_reference(upt, pp, imgs, props, dino)
    # ---- reference, stage by stage (same module, same weights) ----
    with torch.no_grad():
        sizes = torch.as_tensor([[224, 224]] * case["B"])
        rp = upt.prepare_region_proposals([dict(scores=p["scores"], labels=p["labels"], boxes=p["boxes"]) for p in props])
        r_prior, r_mask = upt.get_prior(rp, sizes, upt.prior_method)
        r_glob, r_local = upt.clip_head.image_encoder(imgs, (r_prior, r_mask))
        import contextlib, io
        with contextlib.redirect_stdout(io.StringIO()):
            r_logits, r_pri, r_bh, r_bo, r_obj, _, _ = upt.compute_roi_embeddings(r_glob, dino, r_local, sizes, rp)
    # region proposals must come back unchanged (NMS-safe synthetic boxes, humans first)
    for a, b in zip(rp, props):
        assert torch.equal(a["boxes"], b["boxes"]) and torch.equal(a["labels"], b["labels"]), "proposals altered by NMS/thresholds"

    # ---- oracle ----
    o_dets, inter = O.hoi_forward(imgs, props, dino, enc, head, return_intermediates=True)

    dev = {}
    dev["prior"] = float((inter["prior"] - r_prior).abs().max())
    assert torch.equal(inter["mask"], r_mask)
    r_tokens_local = r_local.permute(0, 2, 3, 1).reshape(case["B"], 196, 512)
    dev["feat_local"] = float((inter["tokens"][:, 1:] - r_tokens_local).abs().max())
    dev["feat_global"] = float((inter["feat_global"] - r_glob).abs().max())
    def nandev(a, b):
        # a fully-outside box gives a zero feature -> 0/0 = NaN in the reference too (U:1048-1050): NaNs must coincide
        assert torch.equal(torch.isnan(a), torch.isnan(b)), "NaN pattern differs"
        d = (a - b).abs()
        d = d[~torch.isnan(d)]
        return float(d.max()) if d.numel() else 0.0

    dev["logits"] = max(nandev(a, b) for a, b in zip(inter["logits"], r_logits))
    dev["nan_logit_rows"] = int(sum(int(torch.isnan(l).any(dim=1).sum()) for l in r_logits))
    dev["prior_scores"] = max(float((a - b).abs().max()) for a, b in zip(inter["priors"], r_pri))
    dev["scores"] = 0.0
    for od, rd in zip(o_dets, ref_dets):
        assert torch.equal(od["pairing"], rd["pairing"]), "pairing differs"
        assert torch.equal(od["labels"], rd["labels"]), "labels differ"
        assert torch.equal(od["objects"], rd["objects"]), "objects differ"
        assert torch.equal(od["boxes"], rd["boxes"])
        rel = (od["scores"] - rd["scores"]).abs() / rd["scores"].abs().clamp_min(1e-30)
        assert torch.equal(torch.isnan(od["scores"]), torch.isnan(rd["scores"]))
        rel = rel[~torch.isnan(rel)]
        dev["scores"] = max(dev["scores"], float(rel.max()) if rel.numel() else 0.0)
    print(f"[{name}] oracle vs reference: " + ", ".join(f"{k}={v:.2e}" for k, v in dev.items()))
    assert dev["prior"] < 1e-4 and dev["feat_local"] < 2e-3 and dev["logits"] < 1e-3 and dev["scores"] < 1e-4, dev

    # ---- fixture (reference outputs only) ----
    out = dict(feat_global=r_glob.numpy(), num_images=np.int64(case["B"]))
    if not case.get("compact"):
        out.update(prior=r_prior.numpy(), mask=r_mask.numpy(), tokens_local=r_tokens_local.numpy().astype(np.float32))
    for b, (d, lg, pr) in enumerate(zip(ref_dets, r_logits, r_pri)):
        out[f"logits_{b}"] = lg.numpy()
        out[f"pairing_{b}"] = d["pairing"].numpy()
        out[f"scores_{b}"] = d["scores"].numpy()
        out[f"labels_{b}"] = d["labels"].numpy()
        out[f"objects_{b}"] = d["objects"].numpy()
        out[f"boxes_{b}"] = d["boxes"].numpy()
    np.savez_compressed(GOLD / f"{name}.npz", **out)
    return dev


def proposals_case(refs):
    """Golden for prepare_region_proposals (U:1361-1406): overlapping raw detections, min/max-instance logic."""
    upt, _ = refs[(117, "hicodet")]
    g = torch.Generator().manual_seed(77)
    results, out = [], {}
    for b, (nh, no) in enumerate([(1, 2), (30, 40), (6, 9), (0, 5)]):
        n = nh + no
        ctr = torch.rand(n, 2, generator=g) * 180 + 20
        wh = torch.rand(n, 2, generator=g) * 60 + 10
        boxes = torch.cat([ctr - wh / 2, ctr + wh / 2], 1)
        labels = torch.cat([torch.zeros(nh, dtype=torch.int64), torch.randint(1, 80, (no,), generator=g)])
        scores = torch.rand(n, generator=g)
        perm = torch.randperm(n, generator=g)
        results.append(dict(scores=scores[perm], labels=labels[perm], boxes=boxes[perm]))
    upt.min_instances, upt.max_instances, upt.box_score_thresh, upt.human_idx = 3, 15, 0.2, 0
    with torch.no_grad():
        rp = upt.prepare_region_proposals(results)
    mine = O.prepare_region_proposals(results, 0, 0.2, 3, 15)
    for b, (a, m) in enumerate(zip(rp, mine)):
        assert torch.equal(a["boxes"], m["boxes"]) and torch.equal(a["labels"], m["labels"]) and torch.equal(a["scores"], m["scores"])
        out[f"boxes_{b}"] = a["boxes"].numpy()
        out[f"scores_{b}"] = a["scores"].numpy()
        out[f"labels_{b}"] = a["labels"].numpy()
        out[f"in_boxes_{b}"] = results[b]["boxes"].numpy()
        out[f"in_scores_{b}"] = results[b]["scores"].numpy()
        out[f"in_labels_{b}"] = results[b]["labels"].numpy()
    np.savez_compressed(GOLD / "proposals.npz", **out)
    print("[proposals] oracle == reference on", len(rp), "images; kept", [len(a["boxes"]) for a in rp])


def roi_align_case():
    """Pin the RoIAlign restatement against the installed torchvision op (the reference's third-party dependency)."""
    import torchvision
    g = torch.Generator().manual_seed(5)
    feat = torch.randn(14, 14, 512, generator=g)
    boxes = torch.rand(40, 4, generator=g) * 300 - 40
    boxes = torch.stack([boxes[:, :2].min(1).values, boxes[:, 2:].min(1).values,
                         boxes[:, :2].max(1).values + 1, boxes[:, 2:].max(1).values + 1], 1)
    boxes[0] = torch.tensor([0.0, 0.0, 224.0, 224.0])
    boxes[1] = torch.tensor([-100.0, -100.0, -50.0, -40.0])   # fully outside -> all samples zeroed
    boxes[2] = torch.tensor([50.0, 50.0, 50.0, 50.0])         # degenerate
    tv = torchvision.ops.roi_align(feat.permute(2, 0, 1)[None], [boxes], output_size=(7, 7), spatial_scale=14 / 224,
                                   aligned=True).flatten(2).mean(-1)
    mine = torch.from_numpy(O.roi_align_mean(feat.numpy(), boxes.numpy(), 14 / 224))
    dev = float((tv - mine).abs().max())
    print(f"[roi_align] restatement vs torchvision {torchvision.__version__}: max abs {dev:.2e}")
    assert dev < 1e-5
    np.savez_compressed(GOLD / "roi_align.npz", feat=feat.numpy(), boxes=boxes.numpy(), out=tv.numpy())
    return dev


def generator_chain_case():
    """Pin hoigen_b200.synthetic.generated_rows (the restated cache-synthesis chain of main_tip_finetune.py:749-824)
    against the reference's OWN classes: Generator (M:247-261), PromptLearner_hoi.forward (M:106-115), TextEncoder
    (M:262-279) over the reference's vanilla CLIP text transformer, mlp_net (M:313-324).  main_tip_finetune.py cannot be
    imported (missing vcoco_text_label.py, dino/, clipnet/ side imports), so those class definitions are compiled
    straight from its source with `ast` — unmodified — into an empty namespace."""
    import ast
    RH.install_shims()
    import CLIP.clip.model as vanilla
    src = (RH.REF / "main_tip_finetune.py").read_text()
    tree = ast.parse(src)
    want = {"weights_init", "Generator", "TextEncoder", "mlp_net", "PromptLearner_hoi"}
    body = [n for n in tree.body if isinstance(n, (ast.ClassDef, ast.FunctionDef)) and n.name in want]
    assert {n.name for n in body} == want
    ns = {"torch": torch, "nn": torch.nn}
    exec(compile(ast.Module(body=body, type_ignores=[]), "main_tip_finetune.py", "exec"), ns)

    targets = torch.tensor([0, 7, 7, 133, 599, 42, 318, 5])
    rows, inp = S.generated_rows(targets, 600, seed=3, n_ctx=5, return_inputs=True)
    torch.manual_seed(0)
    clip_model = vanilla.CLIP(512, 224, 12, 768, 16, 77, 49408, 512, 8, 12)
    missing, unexpected = clip_model.load_state_dict(inp["text"], strict=False)
    assert not unexpected and all(not k.startswith(("transformer.", "ln_final", "text_projection", "positional_embedding"))
                                  for k in missing)
    gen = ns["Generator"]()
    gen.load_state_dict({k[len("gen."):]: v for k, v in inp["gen"].items() if k.startswith("gen.")})
    mlp = ns["mlp_net"](512, 512, 512)
    mlp.load_state_dict({k[len("mlp."):]: v for k, v in inp["gen"].items() if k.startswith("mlp.")})
    enc = ns["TextEncoder"](clip_model)
    pl = ns["PromptLearner_hoi"].__new__(ns["PromptLearner_hoi"])     # its __init__ needs the BPE tokenizer; forward does not
    torch.nn.Module.__init__(pl)
    pl.ctx = torch.nn.Parameter(inp["gen"]["ctx"])
    pl.register_buffer("token_prefix", inp["table"][:, :1])
    pl.register_buffer("token_suffix", inp["table"][:, 1 + 5:])
    tokenized = torch.zeros(600, 77, dtype=torch.int64)
    tokenized[torch.arange(600), inp["eot"]] = 49407                  # argmax picks the EOT position (M:277)
    with torch.no_grad():
        bias = gen(inp["z"])
        f = enc(pl(bias, targets), tokenized[targets])
        f = f / f.norm(dim=-1, keepdim=True)
        ref_rows = mlp(f)
    dev = float((rows - ref_rows).abs().max())
    print(f"[generator chain] restatement vs reference classes: max abs {dev:.2e} (row norm {float(ref_rows.norm(dim=-1).mean()):.3f})")
    assert dev < 1e-5
    return dev


def main():
    assert RH.available(), "needs /root/reference"
    GOLD.mkdir(parents=True, exist_ok=True)
    torch.set_num_threads(8)
    pin = {"torch": torch.__version__, "cases": {}}
    if (GOLD / "PINNING.json").exists() and sys.argv[1:]:     # a partial run keeps the other cases' records
        pin = json.load(open(GOLD / "PINNING.json"))
        pin["torch"] = torch.__version__
    pin["roi_align_vs_torchvision"] = roi_align_case()
    refs = {}
    only = sys.argv[1:] or list(CASES)
    for name in only:
        pin["cases"][name] = run_case(name, CASES[name], refs)
    base = (117, "hicodet", ())
    if base in refs:
        proposals_case({(117, "hicodet"): refs[base]})
    if "hico117_uc0_n16384_b2" in only:
        pin["generator_chain_vs_reference"] = generator_chain_case()
    with open(GOLD / "PINNING.json", "w") as f:
        json.dump(pin, f, indent=1)
    print("wrote", sorted(p.name for p in GOLD.iterdir()))


if __name__ == "__main__":
    main()
