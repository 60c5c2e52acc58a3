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
