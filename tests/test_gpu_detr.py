"""f3, second half: the DETR detector of the proposal stage (U:1594-1599) on the repo's kernels (hoigen_b200/detr.py).
Kernel-level checks against torch, the whole head against the golden made with the UNMODIFIED reference DETR classes
(oracle/make_golden_detr.py), the whole detector (backbone + head) against the stock fp32 modules."""
import math

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def cuda_device():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    return torch.device("cuda:0")


@pytest.mark.parametrize("rows,with_delta,with_norm,pos_rows", [(37, True, True, 37), (800, True, True, 100), (9, False, False, 9),
                                                                 (100, False, True, 0)])
def test_add_layernorm256(cuda_device, rows, with_delta, with_norm, pos_rows):
    """x += delta ; LayerNorm(256) ; side outputs bf16(x), bf16(x + pos[row % pos_rows])  (transformer.py:143-149, 124-125)."""
    from hoigen_b200 import _cabi
    _cabi.init(cuda_device)
    torch.manual_seed(rows)
    x = torch.randn(rows, 256, device=cuda_device) * 2 + 0.3
    delta = torch.randn(rows, 256, device=cuda_device).to(torch.bfloat16) if with_delta else None
    gamma, beta = (torch.rand(256, device=cuda_device) + 0.5, torch.randn(256, device=cuda_device) * 0.1) if with_norm else (None, None)
    pos = torch.randn(pos_rows, 256, device=cuda_device) if pos_rows else None
    ref = x + (delta.float() if with_delta else 0)
    if with_norm:
        ref = torch.nn.functional.layer_norm(ref, (256,), gamma, beta, 1e-5)
    xb = torch.empty(rows, 256, device=cuda_device, dtype=torch.bfloat16)
    xpb = torch.empty_like(xb) if pos_rows else None
    p = lambda t: t.data_ptr() if t is not None else None
    _cabi.call("hoigen_add_layernorm256", x.data_ptr(), p(delta), p(gamma), p(beta), p(pos), pos_rows, xb.data_ptr(), p(xpb), rows)
    assert (x - ref).abs().max().item() <= 2e-5 * max(1.0, ref.abs().max().item())
    assert torch.equal(xb, x.to(torch.bfloat16))
    if pos_rows:
        idx = torch.arange(rows, device=cuda_device) % pos_rows
        assert torch.equal(xpb, (x + pos[idx]).to(torch.bfloat16))


@pytest.mark.parametrize("B,lq,lk,masked", [(2, 100, 100, False), (3, 35, 35, True), (2, 100, 1080, True), (1, 300, 77, True),
                                            (2, 129, 64, False)])
def test_attention_heads32(cuda_device, B, lq, lk, masked):
    """nn.MultiheadAttention's core for 8 heads of 32 with a key-padding mask, strided q / k (packed [q | k] rows), any lengths:
    against torch scaled_dot_product_attention in fp32 on the same bf16 operands (<= 1 bf16 ulp of the output scale), for the
    tensor-core kernel (probabilities rounded to bf16 before P V) and the fp32 SIMT form."""
    from hoigen_b200 import _cabi
    _cabi.init(cuda_device)
    torch.manual_seed(lq * 7 + lk)
    H, D = 8, 256
    qk_q = torch.randn(B * lq, 2 * D, device=cuda_device).to(torch.bfloat16)      # q in columns [0, 256) of a 512-wide row
    kk = torch.randn(B * lk, 2 * D, device=cuda_device).to(torch.bfloat16)        # k in columns [256, 512)
    v = torch.randn(B * lk, D, device=cuda_device).to(torch.bfloat16)
    mask = None
    if masked:
        mask = torch.rand(B, lk, device=cuda_device) < 0.3
        mask[:, 0] = False                                                         # never a fully masked row
        mask_u8 = mask.to(torch.uint8).contiguous()
    out = torch.full((B * lq, D), 7.0, device=cuda_device, dtype=torch.bfloat16)
    q_view, k_view = qk_q[:, :D], kk[:, D:]
    _cabi.call("hoigen_attention_heads32", q_view.data_ptr(), 2 * D, k_view.data_ptr(), 2 * D, v.data_ptr(), D, out.data_ptr(), D,
               mask_u8.data_ptr() if masked else None, B, lq, lk, H, 1.0 / math.sqrt(32))
    qf = q_view.float().reshape(B, lq, H, 32).transpose(1, 2)
    kf = k_view.float().reshape(B, lk, H, 32).transpose(1, 2)
    vf = v.float().reshape(B, lk, H, 32).transpose(1, 2)
    am = None if mask is None else (~mask)[:, None, None, :]
    ref = torch.nn.functional.scaled_dot_product_attention(qf, kf, vf, attn_mask=am).transpose(1, 2).reshape(B * lq, D)
    err = (out.float() - ref).abs().max().item()
    assert err <= 2 ** -7 * max(1.0, ref.abs().max().item()), err


def _seeded_detr(device, classes=91):
    from oracle import detr_ref as D
    det = D.DetrRef(num_classes=classes).eval()
    D.seeded_state(det, 17)
    return det.to(device)


def test_detr_head_matches_reference_golden(cuda_device):
    """Position encoding + input projection + 6 + 6-layer transformer + heads on the kernels against the outputs of the UNMODIFIED
    reference classes (tests/golden/detr_head.npz: models/transformer.py Transformer, PositionEmbeddingSine, MLP evaluated as
    U:1595-1599 on a padded 2-image batch, seeded weights).  bf16 operands through 12 layers: logits within 5e-2 (|logits| <= 2.6),
    boxes within 1e-2, and the oracle restatement reproduces the golden to 1e-4 on the same GPU."""
    from hoigen_b200.detr import KernelDetr
    gold = np.load("tests/golden/detr_head.npz")
    det = _seeded_detr(cuda_device, int(gold["classes"]))
    src = torch.from_numpy(gold["src"].astype(np.float32)).to(cuda_device)
    mask = torch.from_numpy(gold["mask"]).to(cuda_device)
    g_logits, g_boxes = torch.from_numpy(gold["logits"]).to(cuda_device), torch.from_numpy(gold["boxes"]).to(cuda_device)
    o_logits, o_boxes = det.forward_features(src, mask)
    assert (o_logits - g_logits).abs().max().item() < 1e-3 and (o_boxes - g_boxes).abs().max().item() < 1e-4
    fast = KernelDetr(det, with_backbone=False)
    for rep in range(2):
        logits, boxes = fast.head_from_features(src, mask)
        assert logits.shape == g_logits.shape and boxes.shape == g_boxes.shape
        el, eb = (logits - g_logits).abs().max().item(), (boxes - g_boxes).abs().max().item()
        assert el < 5e-2 and eb < 6e-3, (el, eb)
    top2 = g_logits.topk(2, dim=-1).values
    clear = (top2[..., 0] - top2[..., 1]) > 2 * el           # queries whose winning class is not a near-tie in the reference
    agree = (logits.argmax(-1) == g_logits.argmax(-1))
    print(f"DETR head vs reference golden: logits max-abs {el:.3e} (|ref| <= {g_logits.abs().max().item():.2f}), boxes {eb:.3e}, "
          f"argmax agreement {agree.float().mean().item():.3f} overall, {int(clear.sum())} queries with a clear winner")
    assert agree[clear].all()


def test_detr_full_detector_matches_stock_modules(cuda_device):
    """The whole opt-in detector (KernelDetr.forward: FrozenBatchNorm ResNet-50 body on the convolution plan, haloed rows used as
    tokens with the halo as padding keys, transformer, heads) on a padded two-image batch against the stock fp32 modules run the
    way U:1593-1599 runs them."""
    import torchvision
    from torchvision.models._utils import IntermediateLayerGetter
    from torchvision.ops.misc import FrozenBatchNorm2d
    from hoigen_b200.detr import KernelDetr
    from oracle import detr_ref as D
    torch.manual_seed(5)
    det = _seeded_detr(cuda_device)
    r50 = torchvision.models.resnet50(weights=None, norm_layer=FrozenBatchNorm2d)

    class BackboneBase(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.body = IntermediateLayerGetter(r50, return_layers={"layer4": "0"})

    det.backbone = torch.nn.Sequential(BackboneBase(), torch.nn.Identity()).to(cuda_device).eval()
    B, Hh, Ww = 2, 224, 288
    images = torch.randn(B, 3, Hh, Ww, device=cuda_device)
    mask = torch.zeros(B, Hh, Ww, dtype=torch.bool, device=cuda_device)
    images[1, :, 160:, :] = 0; images[1, :, :, 200:] = 0                      # image 1 is 160 x 200, zero-padded (NestedTensor)
    mask[1, 160:, :] = True; mask[1, :, 200:] = True
    with torch.no_grad():
        feat = det.backbone[0].body(images)["0"]
        m = torch.nn.functional.interpolate(mask[None].float(), size=feat.shape[-2:]).to(torch.bool)[0]
        ref_logits, ref_boxes = det.forward_features(feat, m)
    fast = KernelDetr(det)
    logits, boxes = fast(images, mask)
    el, eb = (logits - ref_logits).abs().max().item(), (boxes - ref_boxes).abs().max().item()
    print(f"full DETR on kernels vs stock fp32: logits max-abs {el:.3e} (|ref| <= {ref_logits.abs().max().item():.2f}), boxes {eb:.3e}")
    assert el < 1e-1 and eb < 1e-2, (el, eb)


def test_upt_forward_with_accelerated_detr(cuda_device):
    """UPT.forward (U:1543-1664) with `accelerate_detr()`: the padded NestedTensor batch goes through KernelDetr instead of
    detector.backbone / transformer / heads; what reaches the post-processor (pred_logits, pred_boxes of the last decoder layer)
    matches the stock fp32 modules' run of the same forward, and the detections come out through the normal path."""
    import torchvision
    from torchvision.models._utils import IntermediateLayerGetter
    from torchvision.ops.misc import FrozenBatchNorm2d
    from hoigen_b200 import synthetic as S
    from hoigen_b200.detector import _NestedTensor
    from hoigen_b200.detr import sine_position_embedding
    from test_gpu_e2e import _build
    m, enc, head = _build(117, 256, cuda_device)
    torch.manual_seed(6)
    det = _seeded_detr(cuda_device)
    r50 = torchvision.models.resnet50(weights=None, norm_layer=FrozenBatchNorm2d)

    class BackboneBase(torch.nn.Module):                         # detr/models/backbone.py:60-80
        def __init__(self):
            super().__init__()
            self.body = IntermediateLayerGetter(r50, return_layers={"layer4": "0"})

        def forward(self, nested):
            x = self.body(nested.tensors)["0"]
            mk = torch.nn.functional.interpolate(nested.mask[None].float(), size=x.shape[-2:]).to(torch.bool)[0]
            return {"0": _NestedTensor(x, mk)}

    class Joiner(torch.nn.Sequential):                           # detr/models/backbone.py:94-107
        def forward(self, nested):
            xs = self[0](nested)
            out = list(xs.values())
            return out, [sine_position_embedding(x.mask).permute(0, 3, 1, 2) for x in out]

    det.backbone = Joiner(BackboneBase(), torch.nn.Identity()).to(cuda_device).eval()
    gold = np.load("tests/golden/proposals.npz")
    results = [dict(scores=torch.from_numpy(gold[f"in_scores_{b}"]).to(cuda_device), labels=torch.from_numpy(gold[f"in_labels_{b}"]).to(cuda_device),
                    boxes=torch.from_numpy(gold[f"in_boxes_{b}"]).to(cuda_device)) for b in range(2)]
    seen = []

    class PP(torch.nn.Module):
        def forward(self, outputs, sizes):
            seen.append({k: v.clone() for k, v in outputs.items()})
            return results

    m.detector, m.postprocessor = det, PP()
    imgs = S.make_images(2, seed=7).to(cuda_device)
    dino = S.make_dino_features(2).to(cuda_device)
    m.dino_model = lambda x: dino
    batch = [(torch.randn(3, 200, 264, device=cuda_device), imgs[0]), (torch.randn(3, 168, 224, device=cuda_device), imgs[1])]
    dets_stock = m(batch)
    m.accelerate_detr()
    dets_fast = m(batch)
    assert len(seen) == 2 and seen[0]["pred_logits"].shape == seen[1]["pred_logits"].shape == (2, 100, 92)
    el = (seen[0]["pred_logits"] - seen[1]["pred_logits"]).abs().max().item()
    eb = (seen[0]["pred_boxes"] - seen[1]["pred_boxes"]).abs().max().item()
    print(f"UPT.forward, accelerated DETR vs stock: pred_logits max-abs {el:.3e}, pred_boxes {eb:.3e}")
    assert 0 < el < 1e-1 and eb < 1e-2, (el, eb)
    for a, b in zip(dets_stock, dets_fast):                       # same (stubbed) proposals -> identical detections
        for k in ("pairing", "labels", "objects", "scores"):
            assert torch.equal(a[k], b[k]), k
