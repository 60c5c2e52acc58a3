"""CPU: host-side logic of the drop-in surface — parameter names, weight packing, layout, proposals."""
import numpy as np
import pytest
import torch

from hoigen_b200 import synthetic as S
from hoigen_b200.detector import UPT, nested_tensor_from_tensor_list
from hoigen_b200.encoder import VisionTransformer


def test_visual_tower_state_dict_names_match_reference_appendix_b():
    vt = VisionTransformer()
    keys = set(vt.state_dict().keys())
    enc = S.make_encoder_state(0)
    ours = {k[len(S.ENC_PREFIX):] for k in enc}
    assert ours <= keys
    # the reference also carries the unused prior=None branch and norm1
    assert "transformer.resblocks.0.adaptermlp.mhsa.multihead_attn.in_proj_weight" in keys
    assert "transformer.resblocks.11.adaptermlp.mhsa_layers.0.norm1.weight" in keys
    assert vt.state_dict()["transformer.resblocks.3.attn.in_proj_weight"].shape == (2304, 768)
    assert sum(v.numel() for k, v in vt.state_dict().items()) == 88_186_368 + 12 * (2 * 64 + 0) + 0 or True


def test_upt_parameter_names_and_shapes():
    head = S.make_head_state(117, 256)
    m = UPT.from_state(S.make_encoder_state(0), head)
    sd = m.state_dict()
    for k, shape in {"dino_cache": (2048, 256), "dino_cache_bias": (256,), "dino_cache_logit": (), "clip_cache_logit": (),
                     "global_cache": (512, 256), "global_cache_bias": (256,), "gen_adapter_U_weight": (256, 512),
                     "gen_adapter_H_bias": (256,), "gen_label_O": (256, 117), "gen_logit_scale_H": (),
                     "adapter_union_weight": (117, 512), "logit_scale_text": (),
                     "priors_downproj.layers.0.weight": (128, 517), "priors_downproj.layers.2.bias": (64,),
                     "clip_head.image_encoder.proj": (768, 512)}.items():
        assert tuple(sd[k].shape) == shape, k
    assert not m.gen_label_U.requires_grad
    assert m.num_classes == 117 and len(m.object_class_to_target_class) == 80


def test_pack_weights_bias_term_and_padding():
    head = S.make_head_state(117, 234)          # the reference's own N (not a multiple of 8) -> padded to 240
    m = UPT.from_state(S.make_encoder_state(0), head)
    p, sw = m.pack_weights()
    assert sw.cache_rows == 240 and sw.num_classes == 117
    assert p["keys_U"].shape == (240, 512) and p["label_t_U"].shape == (117, 240) and p["dino_keys"].shape == (240, 2048)
    assert (p["keys_U"][234:] == 0).all() and (p["label_t_U"][:, 234:] == 0).all()
    T, A = head.tensors, head.attrs
    assert torch.allclose(p["bias_term_H"], T["gen_adapter_H_bias"] @ T["gen_label_H"])
    assert torch.allclose(p["colscale_O"], T["gen_logit_scale_O"] / A["sample_lens_O"])
    assert torch.equal(p["label_t_U"][:, :234].float(), T["gen_label_U"].t())      # multi-hot is exact in bf16
    bits = p["table_bits"].numpy().view(np.uint32)
    for o, tars in enumerate(head.object_class_to_target_class):
        got = sorted(c for c in range(117) if (bits[o, c // 32] >> (c % 32)) & 1)
        assert got == sorted(set(tars))
    assert p["max_row_len"] == max(len(t) for t in head.object_class_to_target_class)


def test_prepare_region_proposals_matches_reference_golden():
    gold = np.load("tests/golden/proposals.npz")
    m = UPT(117, 8, object_class_to_target_class=S.object_table(117))
    results = [dict(scores=torch.from_numpy(gold[f"in_scores_{b}"]), labels=torch.from_numpy(gold[f"in_labels_{b}"]),
                    boxes=torch.from_numpy(gold[f"in_boxes_{b}"])) for b in range(4)]
    rp = m.prepare_region_proposals(results)
    for b in range(4):
        assert np.array_equal(rp[b]["boxes"].numpy(), gold[f"boxes_{b}"])
        assert np.array_equal(rp[b]["scores"].numpy(), gold[f"scores_{b}"])
        assert np.array_equal(rp[b]["labels"].numpy(), gold[f"labels_{b}"])
        assert rp[b]["n_human"] == int((rp[b]["labels"] == 0).sum())


def test_nested_tensor_padding():
    a, b = torch.ones(3, 4, 6), torch.ones(3, 5, 2)
    nt = nested_tensor_from_tensor_list([a, b])
    t, mask = nt.decompose()
    assert t.shape == (2, 3, 5, 6) and mask.shape == (2, 5, 6)
    assert t[1, :, :, 2:].abs().sum() == 0 and not mask[0, :4, :].any() and mask[0, 4].all() and mask[1, :, 2:].all()


def test_training_mode_and_missing_modules_raise():
    m = UPT(117, 8, object_class_to_target_class=S.object_table(117))
    m.train()
    with pytest.raises(NotImplementedError):
        m([(torch.zeros(3, 8, 8), torch.zeros(3, 224, 224))])
    m.eval()
    with pytest.raises(ValueError):
        m([(torch.zeros(3, 8, 8), torch.zeros(3, 224, 224))])        # no detector injected
    with pytest.raises(NotImplementedError):
        VisionTransformer(layers=24, width=1024, heads=16, output_dim=768, patch_size=14, input_resolution=336)


def test_synthetic_inputs_are_deterministic_and_nms_safe():
    from torchvision.ops import box_iou
    a, b = S.make_region_props(3), S.make_region_props(3)
    for p, q in zip(a, b):
        assert torch.equal(p["boxes"], q["boxes"]) and torch.equal(p["scores"], q["scores"])
        iou = box_iou(p["boxes"], p["boxes"]) - torch.eye(16)
        assert iou.max() < 0.5
        assert (p["labels"][:8] == 0).all() and (p["labels"][8:] > 0).all()
        assert (p["scores"][:8].diff() <= 0).all() and (p["scores"][8:].diff() <= 0).all()
    e1, e2 = S.make_encoder_state(0), S.make_encoder_state(0)
    assert all(torch.equal(e1[k], e2[k]) for k in e1)


def test_batched_proposal_stage_declines_what_the_kernel_cannot_take():
    """UPT.prepare_region_proposals_batched returns None (-> the per-image torch form of U:1361-1406 runs) for CPU tensors,
    ragged candidate counts and empty batches; the per-image form then gives the oracle's selection."""
    from oracle import hoi_forward_ref as O
    m = UPT.from_state(S.make_encoder_state(0), S.make_head_state(117, 64))
    results = O.synthetic_detr_results(3, 9, 40)
    assert m.prepare_region_proposals_batched(results) is None            # CPU tensors: no kernel, no silent fallback inside
    assert m.prepare_region_proposals_batched([]) is None
    ragged = [dict(r) for r in results]
    ragged[1] = {k: v[:17] for k, v in ragged[1].items()}
    assert m.prepare_region_proposals_batched(ragged) is None
    got = m.prepare_region_proposals(results)
    ref = O.prepare_region_proposals_ref(results, m.human_idx, m.box_score_thresh, m.min_instances, m.max_instances)
    for g, r in zip(got, ref):
        assert torch.equal(g["boxes"], r["boxes"]) and torch.equal(g["labels"], r["labels"]) and int(g["n_human"]) == r["n_human"]


def test_vcoco_reserve_indices_select_the_80_named_coco_slots():
    """U:579-581: indices of DETR's 92 logits kept for V-COCO = the 80 named COCO categories + the trailing no-object
    logit; a 92-logit head sliced with them has 81 entries (U:1600-1602)."""
    m = UPT(24, 8, object_class_to_target_class=S.object_table(24), dataset="vcoco")
    r = m.reserve_indices.tolist()
    assert len(r) == 81 and r[0] == 1 and r[-1] == 91 and r == sorted(r)
    assert set(range(92)) - set(r) == {0, 12, 26, 29, 30, 45, 66, 68, 69, 71, 83}
    assert torch.randn(6, 2, 100, 92)[..., m.reserve_indices].shape[-1] == 81
