"""CPU, builder container only (needs /root/reference): the drop-in wrapper keeps the reference's state-dict contract."""
import os

import pytest
import torch

REF = "/root/reference/upt_tip_cache_model_free_finetune_distill3.py"


@pytest.mark.skipif(not os.path.exists(REF), reason="reference tree not present (GPU box)")
def test_from_reference_keeps_state_dict_and_attributes():
    import subprocess
    import sys
    # run in a subprocess: the harness chdir()s into the reference tree and monkey-patches .cuda() on CPU
    code = r'''
import sys
sys.path.insert(0, "/root/repo")
import torch
from oracle import ref_harness as RH
from hoigen_b200 import synthetic as S
import hoigen_b200
upt, pp = RH.build_reference_upt(117)
RH.load_synthetic_state(upt, S.make_encoder_state(0), S.make_head_state(117, 256))
ref_keys = set(upt.state_dict().keys())
ref_sd = {k: v.clone() for k, v in upt.state_dict().items()}
ours = hoigen_b200.from_reference(upt)
our_sd = ours.state_dict()
assert set(our_sd.keys()) == ref_keys, (sorted(ref_keys - set(our_sd))[:5], sorted(set(our_sd) - ref_keys)[:5])
for k in ("clip_head.image_encoder.proj", "gen_adapter_U_weight", "priors_downproj.layers.1.weight", "dino_cache",
          "clip_head.image_encoder.transformer.resblocks.7.adaptermlp.mhsa_layers.0.norm3.bias"):
    assert torch.equal(our_sd[k], ref_sd[k]), k
assert ours.num_classes == 117 and ours.detector is upt.detector and ours.hyper_lambda == 2.8
assert torch.equal(ours.sample_lens_U, upt.sample_lens_U) and torch.equal(ours.object_embedding, upt.object_embedding)
ours.load_state_dict(ref_sd, strict=True)          # a reference checkpoint loads strictly
assert torch.equal(ours.reserve_indices, torch.as_tensor(upt.reserve_indices))      # U:579-581 (V-COCO 92-logit DETR heads)
p, sw = ours.pack_weights()
assert sw.num_classes == 117 and sw.cache_rows == 256
print("OK")
'''
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "OK" in r.stdout, r.stderr[-2000:]


@pytest.mark.skipif(not os.path.exists(REF), reason="reference tree not present (GPU box)")
def test_reference_detr_has_the_structure_kernel_detr_reads():
    """f3: the UNMODIFIED reference's DETR object (detr/models/detr.py::build) exposes exactly what hoigen_b200.detr.KernelDetr
    and dino.KernelDetrBackboneBody read — transformer.d_model / nhead, post-norm layers without an encoder norm, the packed
    in_proj weights, input_proj / query_embed / class_embed / bbox_embed.layers, and backbone[0].body as an IntermediateLayerGetter
    over FrozenBatchNorm2d bottlenecks with stride on conv2 — and its state dict (backbone aside) loads STRICTLY into the oracle's
    DetrRef, i.e. the oracle pinned in tests/golden/detr_head.npz and the product read the same parameters."""
    import subprocess
    import sys
    code = r'''
import sys, argparse
sys.path.insert(0, "/root/repo")
import torch
from oracle import ref_harness as RH
RH.install_shims(force_cpu=True)
sys.path.insert(0, "/root/reference/detr")
import models.backbone as BB
BB.is_main_process = lambda: False            # no ImageNet download (backbone.py:90)
from models.detr import build
args = argparse.Namespace(dataset="hicodet", pretrained="", dataset_file="coco", device="cpu", hidden_dim=256, position_embedding="sine",
                          lr_backbone=0.0, masks=False, backbone="resnet50", dilation=False, dropout=0.1, nheads=8, dim_feedforward=2048,
                          enc_layers=6, dec_layers=6, pre_norm=False, num_queries=100, aux_loss=False, set_cost_class=1, set_cost_bbox=5,
                          set_cost_giou=2, bbox_loss_coef=5, giou_loss_coef=2, eos_coef=0.1)
det = build(args)[0]
tr = det.transformer
assert (tr.d_model, tr.nhead) == (256, 8) and tr.encoder.norm is None and not tr.encoder.layers[0].normalize_before
assert len(tr.encoder.layers) == 6 and len(tr.decoder.layers) == 6
e, d = tr.encoder.layers[0], tr.decoder.layers[0]
assert tuple(e.self_attn.in_proj_weight.shape) == (768, 256) and tuple(e.linear1.weight.shape) == (2048, 256)
assert tuple(d.multihead_attn.in_proj_weight.shape) == (768, 256) and hasattr(d, "norm3") and hasattr(tr.decoder, "norm")
assert tuple(det.input_proj.weight.shape) == (256, 2048, 1, 1) and tuple(det.query_embed.weight.shape) == (100, 256)
assert [tuple(l.weight.shape) for l in det.bbox_embed.layers] == [(256, 256), (256, 256), (4, 256)]
body = det.backbone[0].body
assert type(body).__name__ == "IntermediateLayerGetter" and type(body.bn1).__name__ == "FrozenBatchNorm2d"
assert body.layer2[0].conv2.stride == (2, 2) and body.layer2[0].conv1.stride == (1, 1) and body.layer4[0].conv2.dilation == (1, 1)
from hoigen_b200.dino import fold_batchnorms
folded = fold_batchnorms(body)                # FrozenBatchNorm2d (no .eps attribute) folds with its 1e-5
x = torch.randn(1, 3, 64, 96)
with torch.no_grad():
    a = body(x)["0"]
    b = folded.layer4(folded.layer3(folded.layer2(folded.layer1(folded.maxpool(folded.relu(folded.conv1(x)))))))
assert (a - b).abs().max().item() < 1e-4 * max(1.0, a.abs().max().item())
from oracle import detr_ref as D
mine = D.DetrRef(num_classes=det.class_embed.weight.shape[0] - 1)
res = mine.load_state_dict({k: v for k, v in det.state_dict().items() if not k.startswith("backbone.")}, strict=True)
print("OK", res)
'''
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "OK" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]
