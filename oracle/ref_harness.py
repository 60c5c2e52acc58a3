"""Harness around the UNMODIFIED reference (from /root/reference, or its staged byte-identical copy oracle/_ref on the GPU
box): builds its UPT through the real `build_detector` (upt_tip_cache_model_free_finetune_distill3.py:1712), then loads
the seeded synthetic state of hoigen_b200/synthetic.py into it.  Used by oracle/make_golden.py to pin the oracle and
generate tests/golden/*, and by bench.py's reference legs (CPU arm, same-GPU torch comparator).  Nothing under
hoigen_b200/ imports it.

Shims (none alters arithmetic):
  1. sys.path wiring + cwd=/root/reference (the modules do sys.path.append('detr'), `import clip`, `import pocket`)
  2. stub module `ftfy` (tokenizer import only)
  3. stub module `transformer_module` (C:14 imports a file that is not in the repo; both names are shadowed at C:27)
  4. `.cuda()` -> identity on this CPU-only box (the hot path hard-codes .cuda(): U:976-979, 1112-1114, 1133-1135)
  5. detr backbone `is_main_process -> False` (no ImageNet download), non-existent args.pretrained
  6. `torch.load(clip_model_path)` returns an object whose .state_dict() is a random-init vanilla CLIP ViT-B/16
"""
from __future__ import annotations

import argparse
import contextlib
import io
import os
import pickle
import sys
import tempfile
import types
from pathlib import Path

import numpy as np
import torch

def _resolve_ref() -> Path:
    """/root/reference in the builder container; on the GPU box the byte-identical staged copy oracle/_ref/ (made by
    oracle/stage_ref.py from __graft_entry__.build(), git-ignored, travels with the snapshot)."""
    if os.environ.get("HOIGEN_REFERENCE"):
        return Path(os.environ["HOIGEN_REFERENCE"])
    for cand in (Path("/root/reference"), Path(__file__).resolve().parent / "_ref"):
        if (cand / "upt_tip_cache_model_free_finetune_distill3.py").exists():
            return cand
    return Path("/root/reference")


REF = _resolve_ref()


def available() -> bool:
    return (REF / "upt_tip_cache_model_free_finetune_distill3.py").exists()


_installed = False


def install_shims(force_cpu: bool = False):
    """force_cpu: neutralise the reference's hard-coded `.cuda()` calls even when a GPU is present (the CPU reference arm
    of bench.py on the GPU box).  Process-wide monkey-patch: only ever done in a process that runs nothing else."""
    global _installed
    if _installed:
        return
    shim_dir = tempfile.mkdtemp(prefix="hoigen_shims_")
    (Path(shim_dir) / "ftfy.py").write_text("def fix_text(s):\n    return s\n")
    (Path(shim_dir) / "transformer_module.py").write_text(
        "class TransformerDecoderLayer:  # shadowed at CLIP_models_adapter_prior2.py:27\n    pass\n"
        "class TransformerDecoderLayer_womhsa:\n    pass\n")
    for p in [str(REF / "pocket"), str(REF / "CLIP"), str(REF / "detr"), str(REF), shim_dir]:
        sys.path.insert(0, p)
    os.chdir(REF)
    if force_cpu or not torch.cuda.is_available():
        torch.nn.Module.cuda = lambda self, *a, **k: self
        torch.Tensor.cuda = lambda self, *a, **k: self
    _installed = True


def default_args(**over) -> argparse.Namespace:
    """main_tip_finetune.py:1048-1194 defaults + its hard-sets (M:393-396, 444-445, 834) for the --eval path."""
    d = dict(
        backbone="resnet50", dilation=False, position_embedding="sine", repr_dim=512, hidden_dim=256, enc_layers=6,
        dec_layers=6, dim_feedforward=2048, dropout=0.1, nheads=8, num_queries=100, pre_norm=False, aux_loss=True,
        set_cost_class=1, set_cost_bbox=5, set_cost_giou=2, bbox_loss_coef=5, giou_loss_coef=2, eos_coef=0.1,
        alpha=0.5, gamma=0.2, dataset="hicodet", device="cpu", pretrained="/nonexistent/detr.pth", eval=True, cache=False,
        box_score_thresh=0.2, fg_iou_thresh=0.5, min_instances=3, max_instances=15, use_insadapter=True,
        use_distill=False, use_consistloss=False, logits_type="HO+U+T", num_shot=2, file1="", prior_type="cbe",
        obj_affordance=False, zs=False, hyper_lambda=2.8, use_weight_pred=False, zs_type="rare_first",
        fill_zs_verb_type=0, pseudo_label=False, tpt=False, vis_tor=1.0, adapter_num_layers=1, N_CTX=24, CSC=False,
        CTX_INIT="", CLASS_TOKEN_POSITION="end", use_templates=False, LA=False, LA_weight=0.6, feat_mask_type=0,
        num_classes=117, prior_method=0, vis_prompt_num=50, box_proj=0, adapter_pos="all", use_multi_hot=True,
        label_learning=False, label_choice="random", use_mlp_proj=False, masks=False, frozen_weights=None,
        lr_backbone=0.0, keep_datasets=10 ** 9, human_idx=0,
        # hard-set in main (M:393-396, 444-445)
        dino=True, clip_global=True, cache_model="gen_feat", generate_feature=True,
    )
    d.update(over)
    return argparse.Namespace(**d)


def _synthetic_cache_pickle(path: Path, num_classes: int, table, rng: np.random.Generator):
    """`file1` schema read by load_cache_model (U:636-688). Every class gets >= num_shot samples."""
    anno = {}
    img = 0
    if num_classes in (117, 24):
        for obj, verbs in enumerate(table):
            for v in verbs:
                for _ in range(2):
                    P = 1
                    anno[f"img_{img:06d}.jpg"] = {
                        "verbs": np.array([v]), "objects": np.array([obj]),
                        "boxes_h": rng.uniform(0, 100, (P, 4)).astype(np.float32) + np.array([0, 0, 100, 100], np.float32),
                        "boxes_o": rng.uniform(0, 100, (P, 4)).astype(np.float32) + np.array([0, 0, 100, 100], np.float32),
                        "union_features": rng.standard_normal((P, 512)).astype(np.float32),
                        "object_features": rng.standard_normal((P, 512)).astype(np.float32),
                        "huamn_features": rng.standard_normal((P, 512)).astype(np.float32),   # sic (U:686)
                    }
                    img += 1
    elif num_classes == 600:
        # HOI classes: (object, verb) pairs from the correspondence table; load_cache_model maps them to interaction ids
        # through object_n_verb_to_interaction (U:650-651)
        from hoigen_b200 import synthetic as S
        for _hoi, obj, v in S.load_object_tables()["hico_correspondence"]:
            for _ in range(2):
                P = 1
                anno[f"img_{img:06d}.jpg"] = {
                    "verbs": np.array([v]), "objects": np.array([obj]),
                    "boxes_h": rng.uniform(0, 100, (P, 4)).astype(np.float32) + np.array([0, 0, 100, 100], np.float32),
                    "boxes_o": rng.uniform(0, 100, (P, 4)).astype(np.float32) + np.array([0, 0, 100, 100], np.float32),
                    "union_features": rng.standard_normal((P, 512)).astype(np.float32),
                    "object_features": rng.standard_normal((P, 512)).astype(np.float32),
                    "huamn_features": rng.standard_normal((P, 512)).astype(np.float32),   # sic (U:686)
                }
                img += 1
    else:
        raise NotImplementedError(num_classes)
    with open(path, "wb") as f:
        pickle.dump(anno, f)


class _StubDetr(torch.nn.Module):
    """Exposes what UPT.forward touches on `detector` (U:1594-1599) without running DETR."""

    def __init__(self):
        super().__init__()
        self.query_embed = torch.nn.Embedding(1, 1)
        self.class_embed = lambda hs: hs
        self.bbox_embed = lambda hs: hs
        self.input_proj = lambda x: x

    def backbone(self, nested):
        from detr.util.misc import NestedTensor
        t = nested.tensors
        return [NestedTensor(t[:, :1, :1, :1], torch.zeros(t.shape[0], 1, 1, dtype=torch.bool, device=t.device))], [None]

    def transformer(self, src, mask, query, pos):
        return torch.zeros(1, src.shape[0], 1, 4, device=src.device), None


class _StubPostprocessor(torch.nn.Module):
    """Feeds synthetic boxes through the REAL prepare_region_proposals (U:1361)."""

    def __init__(self):
        super().__init__()
        self.results = None

    def forward(self, outputs, sizes):
        return self.results


def build_reference_upt(num_classes: int = 117, dataset: str = "hicodet", quiet: bool = True, force_cpu: bool = False,
                        **arg_over):
    """Real build_detector on CPU -> (upt, stub_postprocessor)."""
    install_shims(force_cpu)
    import detr.models.backbone as backbone_mod
    backbone_mod.is_main_process = lambda: False
    import CLIP.clip.model as vanilla
    import upt_tip_cache_model_free_finetune_distill3 as U
    from hoigen_b200 import synthetic as S

    args = default_args(num_classes=num_classes, dataset=dataset, **arg_over)
    table = S.object_table(num_classes)
    tmp = Path(tempfile.mkdtemp(prefix="hoigen_ref_"))
    file1 = tmp / ("hico_cache.p" if dataset == "hicodet" else "vcoco_cache.p")
    _synthetic_cache_pickle(file1, num_classes, table, np.random.default_rng(7))
    args.file1 = str(file1)

    torch.manual_seed(0)
    clip_sd = vanilla.CLIP(512, 224, 12, 768, 16, 77, 49408, 512, 8, 12).state_dict()

    class _Holder:
        def state_dict(self):
            return clip_sd

    real_load = torch.load

    def fake_load(path, *a, **k):
        if str(path) == "CLIP_CKPT":
            return _Holder()
        return real_load(path, *a, **k)

    R = 600 if dataset == "hicodet" else 236
    gen_feat = torch.randn(3 * R, 512)
    gen_tar = torch.cat([torch.arange(R)] * 3)
    tables = S.load_object_tables()
    if dataset == "hicodet":
        gen_verb = [c[2] for c in tables["hico_correspondence"]]
        onv = [[(v if v >= 0 else None) for v in row] for row in tables["hico_object_n_verb_to_interaction"]]
    else:
        gen_verb = [0] * R
        onv = [[None] * num_classes for _ in range(81)]

    def _build(n_keys):
        torch.load = fake_load
        try:
            ctx = contextlib.redirect_stdout(io.StringIO()) if quiet else contextlib.nullcontext()
            with ctx:
                return U.build_detector(
                    args, torch.randn(512, n_keys), None, torch.nn.Identity(), torch.randn(2048, n_keys), None,
                    gen_feat, gen_tar, gen_verb, table, table, object_n_verb_to_interaction=onv,
                    clip_model_path="CLIP_CKPT", num_anno=torch.ones(num_classes))
        finally:
            torch.load = real_load

    upt = _build(8)
    stub_pp = _StubPostprocessor()
    upt.detector = _StubDetr()
    upt.postprocessor = stub_pp
    upt.eval()
    return upt, stub_pp


def load_synthetic_state(upt, enc_sd, head):
    """Overwrite every tensor the eval forward reads with the seeded synthetic state (any cache size N)."""
    missing, unexpected = upt.load_state_dict(enc_sd, strict=False)
    assert not unexpected, unexpected
    for name, t in head.tensors.items():
        if name.startswith("priors_downproj."):
            continue
        setattr(upt, name, torch.nn.Parameter(t.clone(), requires_grad=False))
    upt.load_state_dict({k: v for k, v in head.tensors.items() if k.startswith("priors_downproj.")}, strict=False)
    for name, t in head.attrs.items():
        setattr(upt, name, t.clone())
    # U:432,442-445: dino / global cache values are the union labels
    upt.dino_cache_values = head.tensors["gen_label_U"].clone()
    upt.clip_cache_values = head.tensors["gen_label_U"].clone()
    upt.object_class_to_target_class = head.object_class_to_target_class
    upt.hyper_lambda = head.hyper["hyper_lambda"]
    upt.box_score_thresh = head.hyper["box_score_thresh"]
    upt.min_instances = head.hyper["min_instances"]
    upt.max_instances = head.hyper["max_instances"]
    upt.human_idx = head.hyper["human_idx"]
    return upt


class _FixedDino(torch.nn.Module):
    """dino_model stand-in: returns the provided (already normalised) features * 1 so U:1617-1618 renormalises to itself."""

    def __init__(self):
        super().__init__()
        self.feats = None

    def forward(self, x):
        return self.feats.clone()


def run_reference(upt, stub_pp, images: torch.Tensor, detr_results, dino_feats: torch.Tensor, quiet: bool = True):
    """UPT.forward (U:1543) on [(img_detr, img_clip)] with synthetic DETR results injected before NMS."""
    stub_pp.results = [dict(scores=r["scores"], labels=r["labels"], boxes=r["boxes"]) for r in detr_results]
    upt.dino_model = _FixedDino()
    upt.dino_model.feats = dino_feats
    inputs = [(torch.zeros(3, 32, 32, device=images.device), images[b]) for b in range(images.shape[0])]
    ctx = contextlib.redirect_stdout(io.StringIO()) if quiet else contextlib.nullcontext()
    with torch.no_grad(), ctx:
        return upt(inputs)
