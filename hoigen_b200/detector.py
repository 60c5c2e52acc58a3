"""B200 HOI detector behind the reference's `UPT` surface.

Mirrors upt_tip_cache_model_free_finetune_distill3.py: `UPT.forward(images, targets=None)` (U:1543, eval branch),
`prepare_region_proposals` (U:1361), `recover_boxes`, the parameter / attribute names of SURVEY.md Appendix B
(so reference checkpoints load and `net.module.num_classes`, `.object_class_to_target_class` keep working), and
the detection dicts of U:1421-1425.  From the region proposals onward everything runs in sm_100a kernels through
the C ABI: get_prior -> VisionTransformer(x, prior) -> RoIAlign / pair assembly -> cache + text logits ->
prior scores + ordered triplet emission.  There is NO PyTorch fallback for those stages.

Out of scope (delegated to injected stock modules, exactly as the reference composes them): the DETR proposal
network (`detector`, `postprocessor`) and the DINO ResNet-50 (`dino_model`) — SURVEY.md §8 rows a8 / f3.
"""
from __future__ import annotations

import ctypes as C
import math
import os
from typing import Dict, List, Optional, Sequence

import torch
from torch import nn

from . import _cabi
from .encoder import MAX_PRIOR_TOKENS, TOKENS, VisionTransformer, _ParamBag, _linear_params


# slots of DETR's 91-way COCO classifier that carry no category ('N/A' in the reference's COCO_CLASSES list, U:570-578)
_COCO_UNNAMED_SLOTS = (0, 12, 26, 29, 30, 45, 66, 68, 69, 71, 83)


class PackedDetections:
    """The batch's detections as the kernels wrote them: flat per-field tensors + host-side CSR offsets.
    `pairing` holds, per image b, a contiguous [2][M_b] block starting at 2*triplet_off[b]."""

    def __init__(self, scores, labels, objects, pairing, boxes, triplet_off, box_off, size):
        self.scores, self.labels, self.objects, self.pairing, self.boxes = scores, labels, objects, pairing, boxes
        self.triplet_off, self.box_off, self.size = list(triplet_off), list(box_off), tuple(size)

    @property
    def num_images(self) -> int:
        return len(self.triplet_off) - 1

    def image(self, b: int) -> dict:
        s, e = self.triplet_off[b], self.triplet_off[b + 1]
        return dict(boxes=self.boxes[self.box_off[b]: self.box_off[b + 1]], pairing=self.pairing[2 * s: 2 * e].view(2, e - s),
                    scores=self.scores[s:e], labels=self.labels[s:e], objects=self.objects[s:e],
                    size=torch.tensor(self.size, dtype=torch.int64, device=self.scores.device))


class _PinnedSlot:
    buf = None
    event = None
    generation = 0


class _PendingForward:
    """Handle of a forward that has been enqueued but not waited for (UPT.launch_from_proposals -> UPT.finish)."""


class DetectionList(list):
    """List[dict] exactly as the reference returns it (U:1421-1425), plus `.packed` for the zero-copy gather."""
    packed: "PackedDetections" = None


class _NestedTensor:
    """Minimal stand-in for detr.util.misc.NestedTensor (tensors + padding mask), U:1592-1593."""

    def __init__(self, tensors, mask):
        self.tensors, self.mask = tensors, mask

    def decompose(self):
        return self.tensors, self.mask

    def to(self, device):
        return _NestedTensor(self.tensors.to(device), self.mask.to(device) if self.mask is not None else None)


def nested_tensor_from_tensor_list(tensor_list: Sequence[torch.Tensor]) -> _NestedTensor:
    """Zero-pad (C,H,W) images to the batch maximum; mask is True on padding (detr/util/misc.py:307 semantics)."""
    c = tensor_list[0].shape[0]
    h = max(t.shape[1] for t in tensor_list)
    w = max(t.shape[2] for t in tensor_list)
    out = tensor_list[0].new_zeros((len(tensor_list), c, h, w))
    mask = torch.ones((len(tensor_list), h, w), dtype=torch.bool, device=out.device)
    for i, t in enumerate(tensor_list):
        out[i, :, : t.shape[1], : t.shape[2]].copy_(t)
        mask[i, : t.shape[1], : t.shape[2]] = False
    return _NestedTensor(out, mask)


class _ClipHead(nn.Module):
    """Holds `image_encoder` under the reference's name `clip_head.image_encoder` (CustomCLIP, U:208-217)."""

    def __init__(self, image_encoder: VisionTransformer):
        super().__init__()
        self.image_encoder = image_encoder


class _PriorMLP(_ParamBag):
    """Parameter layout of MLP(517,128,64,3) (U:40-52): layers.{0,1,2}.{weight,bias}."""

    def __init__(self, in_dim: int = 517, hidden: int = 128, out_dim: int = 64):
        super().__init__()
        dims = [in_dim, hidden, hidden, out_dim]
        self.layers = nn.ModuleList([_linear_params(dims[i + 1], dims[i]) for i in range(3)])


class UPT(nn.Module):
    """Eval-mode drop-in for the reference UPT (cache_model='gen_feat', logits_type='HO+U+T', prior_type='cbe',
    prior_method=0, use_insadapter=True — the configuration main_tip_finetune.py hard-sets, M:393-396,444-445)."""

    def __init__(self, num_classes: int, cache_rows: int, *, detector: Optional[nn.Module] = None,
                 postprocessor: Optional[nn.Module] = None, dino_model: Optional[nn.Module] = None,
                 clip_head: Optional[nn.Module] = None, human_idx: int = 0, box_score_thresh: float = 0.2,
                 min_instances: int = 3, max_instances: int = 15, hyper_lambda: float = 2.8,
                 object_class_to_target_class: Optional[List[List[int]]] = None, dino: bool = True,
                 clip_global: bool = True, dataset: str = "hicodet", fold_cache: bool = False,
                 scoring_precision: str = "bf16", cache_affinity: str = "linear", cache_beta: float = 10.0):
        super().__init__()
        # "linear" = the reference (phi = f W^T + b, U:1156-1158).  "exp" = the textbook Tip-Adapter exp(beta (f W^T + b)) that
        # north_star names as an option of the fused cache kernel; it is NOT the reference's arithmetic (oracle: affinity='exp')
        if cache_affinity not in ("linear", "exp"):
            raise ValueError("cache_affinity must be 'linear' or 'exp'")
        if cache_affinity == "exp" and (fold_cache or scoring_precision != "bf16" or num_classes > 128):
            raise ValueError("cache_affinity='exp' runs in the fused bf16 cache kernel only (num_classes <= 128, no fold_cache)")
        self.cache_affinity, self.cache_beta = cache_affinity, float(cache_beta)
        # "bf16" (default, the benchmarked path): bf16 operands, fp32 accumulation, logits within 1e-2 of the reference.
        # "fp32": RoI features and every cache / text product in fp32-equivalent arithmetic (3 x bf16 split on the same
        # tcgen05 GEMM, hoigen_score_pairs_fp32): logits within 1e-4 of the reference GIVEN the same encoder features
        if scoring_precision not in ("bf16", "fp32"):
            raise ValueError("scoring_precision must be 'bf16' or 'fp32'")
        if fold_cache and scoring_precision == "fp32":
            raise ValueError("fold_cache is a bf16 shortcut; use scoring_precision='bf16' with it")
        self.scoring_precision = scoring_precision
        # opt-in: contract every (linear) cache with its label matrix at pack time — see hoigen_score_pairs_folded
        self.fold_cache = fold_cache
        C_, N = num_classes, cache_rows
        self.detector, self.postprocessor, self.dino_model = detector, postprocessor, dino_model
        self.clip_head = clip_head if clip_head is not None else _ClipHead(VisionTransformer())
        self.num_classes, self.human_idx = num_classes, human_idx
        self.box_score_thresh, self.min_instances, self.max_instances = box_score_thresh, min_instances, max_instances
        self.hyper_lambda = hyper_lambda
        self.object_class_to_target_class = object_class_to_target_class
        self.dino, self.clip_global, self.dataset = dino, clip_global, dataset
        self.visual_output_dim = 512
        self.priors_initial_dim = 517
        self.logits_type, self.cache_model, self.prior_type, self.prior_method = "HO+U+T", "gen_feat", "cbe", 0
        self.use_insadapter = True
        # U:579-581: the 80 named classes of DETR's 91-slot COCO head (+ the trailing no-object logit); applied to V-COCO
        # models whose DETR emits 92 logits (U:1600-1602)
        self.reserve_indices = torch.as_tensor([i for i in range(91) if i not in _COCO_UNNAMED_SLOTS] + [91])
        ls = math.log(1 / 0.07)
        P = lambda *shape: nn.Parameter(torch.zeros(*shape))
        if dino:
            self.dino_cache = P(2048, N)
            self.dino_cache_bias = nn.Parameter(-torch.ones(N))
            self.dino_cache_logit = nn.Parameter(torch.ones([]) * ls)
        if clip_global:
            self.clip_cache_logit = nn.Parameter(torch.ones([]) * ls)
            self.global_cache = P(512, N)
            self.global_cache_bias = nn.Parameter(-torch.ones(N))
        for X in ("U", "H", "O"):
            setattr(self, f"gen_adapter_{X}_weight", P(N, 512))
            setattr(self, f"gen_adapter_{X}_bias", nn.Parameter(-torch.ones(N)))
            setattr(self, f"gen_label_{X}", nn.Parameter(torch.zeros(N, C_), requires_grad=False))
            setattr(self, f"gen_logit_scale_{X}", nn.Parameter(torch.ones([]) * ls))
        self.adapter_union_weight = P(C_, 512)
        self.logit_scale_text = nn.Parameter(torch.ones([]) * ls)
        self.priors_downproj = _PriorMLP()
        # plain attributes, re-derived from the labels unless given (U:397-405, 436, 450)
        self.sample_lens_H = self.sample_lens_O = self.sample_lens_U = None
        self.dino_sample_len = self.global_sample_len = None
        self.object_embedding = torch.zeros(80, 512)
        self.origin_text_embeddings = None
        self._packed = None
        self._ws: Dict[str, torch.Tensor] = {}
        self.register_load_state_dict_post_hook(lambda module, incompatible: module.invalidate_packed())

    # ------------------------------------------------------------------------------------------------------
    # construction helpers
    # ------------------------------------------------------------------------------------------------------
    @classmethod
    def from_state(cls, enc_state: Dict[str, torch.Tensor], head, **kw) -> "UPT":
        """Build from a `clip_head.image_encoder.*` state dict + a HeadState (hoigen_b200.synthetic)."""
        N = head.tensors["gen_adapter_U_weight"].shape[0]
        m = cls(head.num_classes, N, object_class_to_target_class=head.object_class_to_target_class,
                human_idx=head.hyper["human_idx"], box_score_thresh=head.hyper["box_score_thresh"],
                min_instances=head.hyper["min_instances"], max_instances=head.hyper["max_instances"],
                hyper_lambda=head.hyper["hyper_lambda"], **kw)
        missing, unexpected = m.load_state_dict({**enc_state, **head.tensors}, strict=False)
        assert not unexpected, unexpected
        # only the unused prior=None adapter branch may be missing
        assert all(".adaptermlp.mhsa." in k or ".norm1." in k for k in missing), missing
        for k, v in head.attrs.items():
            setattr(m, k, v.clone())
        return m.eval()

    @classmethod
    def from_reference(cls, ref: nn.Module) -> "UPT":
        """Wrap an already-built reference UPT (after build_detector + load_state_dict, M:865-880): keeps its DETR,
        postprocessor, DINO model and text tower objects, re-hosts the visual tower and every hot-path parameter."""
        sd = ref.state_dict()
        N = sd["gen_adapter_U_weight"].shape[0]
        ref_clip = ref.clip_head
        m = cls(ref.num_classes, N, detector=ref.detector, postprocessor=ref.postprocessor,
                dino_model=getattr(ref, "dino_model", None), human_idx=ref.human_idx,
                box_score_thresh=ref.box_score_thresh, min_instances=ref.min_instances,
                max_instances=ref.max_instances, hyper_lambda=ref.hyper_lambda,
                object_class_to_target_class=ref.object_class_to_target_class, dino=bool(ref.dino),
                clip_global=bool(ref.clip_global), dataset=ref.dataset)
        vt = VisionTransformer()
        vt.load_state_dict(ref_clip.image_encoder.state_dict(), strict=True)
        ref_clip.image_encoder = vt          # text tower / prompt learner stay the reference's own modules
        m.clip_head = ref_clip
        own = {k: v for k, v in sd.items() if not k.startswith(("detector.", "clip_head.", "dino_model."))}
        missing, unexpected = m.load_state_dict(own, strict=False)
        assert not unexpected, unexpected
        for k in ("sample_lens_H", "sample_lens_O", "sample_lens_U", "dino_sample_len", "global_sample_len",
                  "object_embedding", "origin_text_embeddings"):
            if hasattr(ref, k):
                setattr(m, k, getattr(ref, k))
        if hasattr(ref, "reserve_indices"):
            m.reserve_indices = torch.as_tensor(ref.reserve_indices).clone()
        if getattr(ref, "zs_type", None) == "rare_first":
            m.object_class_to_target_class = ref.object_to_verb   # U:821-822
        return m.eval()

    def invalidate_packed(self) -> None:
        self._packed = None

    def accelerate_dino(self, use_graph: bool = True, engine: str = "kernels") -> "UPT":
        """Opt in to a fast execution of the injected DINO ResNet-50 (row a8 of SURVEY.md §8, U:1616-1618).

        engine = "kernels" (default): hoigen_b200.dino.KernelDinoR50 -- BatchNorms folded, NHWC bf16, every convolution on
        the repo's tcgen05 GEMM (3x3 as an implicit GEMM), one C call per batch.  engine = "cudnn": FastDinoR50 -- the same
        folded module left to cuDNN (bf16 channels-last, one CUDA graph per batch size).  Both give the stock fp32 module's
        features to bf16 accuracy; the stock module stays the default because this row is outside the parity-gated path."""
        from .dino import FastDinoR50, KernelDinoR50
        if self.dino_model is None:
            raise ValueError("no dino_model to accelerate")
        if engine == "kernels":
            fast = KernelDinoR50(self.dino_model)
        elif engine == "cudnn":
            fast = FastDinoR50(self.dino_model, use_graph=use_graph)
        else:
            raise ValueError(f"accelerate_dino: engine must be 'kernels' or 'cudnn', got {engine!r}")
        object.__setattr__(self, "_fast_dino", fast)
        return self

    def accelerate_detr_backbone(self) -> "UPT":
        """Opt in: run the injected DETR detector's ResNet-50 backbone body (U:1594, detr/models/backbone.py:72-73; FrozenBatchNorm,
        layer4 only) on this repo's convolution kernels (hoigen_b200.dino.KernelDetrBackboneBody; bf16 activations).  The
        position encoding, the transformer and the heads stay the injected stock modules."""
        from .dino import KernelDetrBackboneBody
        backbone = self.detector.backbone[0]
        if not isinstance(backbone.body, KernelDetrBackboneBody):
            backbone.body = KernelDetrBackboneBody(backbone.body)
        return self

    def accelerate_detr(self) -> "UPT":
        """Opt in: run the whole injected DETR detector of the proposal stage (U:1594-1599: ResNet-50 backbone body, position
        encoding, input projection, 6 + 6-layer transformer, class / box heads) on this repo's kernels
        (hoigen_b200.detr.KernelDetr: bf16 operands, fp32 residual stream; only the last decoder layer is produced, which is
        all the forward reads).  The stock fp32 modules stay the default: bf16 scores can move proposals that sit next to the
        0.2 score / 0.5 NMS thresholds."""
        from .detr import KernelDetr
        object.__setattr__(self, "_fast_detr", KernelDetr(self.detector))
        return self

    def _apply(self, fn, *a, **k):
        self.invalidate_packed()
        self._ws = {}
        out = super()._apply(fn, *a, **k)
        for name in ("sample_lens_H", "sample_lens_O", "sample_lens_U", "dino_sample_len", "global_sample_len",
                     "object_embedding"):
            t = getattr(self, name, None)
            if isinstance(t, torch.Tensor):
                setattr(self, name, fn(t))
        return out

    # ------------------------------------------------------------------------------------------------------
    # weight packing for the scoring / prior / emit kernels (build-time, not timed)
    # ------------------------------------------------------------------------------------------------------
    @torch.no_grad()
    def pack_weights(self):
        dev = self.adapter_union_weight.device
        C_, N0 = self.num_classes, self.gen_adapter_U_weight.shape[0]
        N = (N0 + 7) // 8 * 8   # TMA row strides are multiples of 16 bytes: pad the cache with all-zero rows (no effect)
        padr = lambda t: torch.nn.functional.pad(t, (0, 0, 0, N - N0)) if N != N0 else t      # rows
        padc = lambda t: torch.nn.functional.pad(t, (0, N - N0)) if N != N0 else t            # columns
        bf = lambda t: t.detach().to(device=dev, dtype=torch.bfloat16).contiguous()
        f32 = lambda t: t.detach().to(device=dev, dtype=torch.float32).contiguous()
        p: Dict[str, torch.Tensor] = {}
        sw = _cabi.ScoreWeights()
        sw.num_classes, sw.cache_rows = C_, N

        def lens(name, label):
            t = getattr(self, name, None)
            return f32(t) if t is not None else f32(label.sum(0))

        for i, X in enumerate(("H", "O", "U")):
            W, b, Y = (getattr(self, f"gen_adapter_{X}_weight"), getattr(self, f"gen_adapter_{X}_bias"),
                       getattr(self, f"gen_label_{X}"))
            s = lens(f"sample_lens_{X}", Y)
            p[f"keys_{X}"] = bf(padr(W.detach()))
            p[f"bias_term_{X}"] = f32(b.float() @ Y.float())                       # (b Y), exact fp32 carrier of the bias
            p[f"label_t_{X}"] = bf(padc(Y.detach().t()))                           # 0/1 -> exact in bf16
            p[f"colscale_{X}"] = f32(getattr(self, f"gen_logit_scale_{X}").float() / s)
            sw.cache_keys[i], sw.bias_term[i] = p[f"keys_{X}"].data_ptr(), p[f"bias_term_{X}"].data_ptr()
            sw.label_t[i], sw.colscale[i] = p[f"label_t_{X}"].data_ptr(), p[f"colscale_{X}"].data_ptr()
        Yu = self.gen_label_U.float()                                              # dino/global cache values (U:432,442)
        if self.clip_global:
            p["global_keys"] = bf(padr(self.global_cache.detach().t()))
            p["global_bias_term"] = f32(self.global_cache_bias.float() @ Yu)
            p["colscale_global"] = f32(self.clip_cache_logit.float() / lens("global_sample_len", Yu))
        else:  # branch absent: zero keys contribute nothing
            p["global_keys"] = torch.zeros(N, 512, device=dev, dtype=torch.bfloat16)
            p["global_bias_term"] = torch.zeros(C_, device=dev)
            p["colscale_global"] = torch.zeros(C_, device=dev)
        sw.global_keys, sw.global_bias_term = p["global_keys"].data_ptr(), p["global_bias_term"].data_ptr()
        sw.colscale_global = p["colscale_global"].data_ptr()
        if self.dino:
            p["dino_keys"] = bf(padr(self.dino_cache.detach().t()))
            p["dino_bias_term"] = f32(self.dino_cache_bias.float() @ Yu)
            p["colscale_dino"] = f32(self.dino_cache_logit.float() / lens("dino_sample_len", Yu))
            sw.dino_keys, sw.dino_bias_term = p["dino_keys"].data_ptr(), p["dino_bias_term"].data_ptr()
            sw.colscale_dino = p["colscale_dino"].data_ptr()
        sw.affinity, sw.beta = (1 if self.cache_affinity == "exp" else 0), self.cache_beta
        if self.cache_affinity == "exp":
            for i, X in enumerate(("H", "O", "U")):
                p[f"cache_bias_{X}"] = f32(padc(getattr(self, f"gen_adapter_{X}_bias").detach()))
                sw.cache_bias[i] = p[f"cache_bias_{X}"].data_ptr()
            if self.clip_global:
                p["global_bias"] = f32(padc(self.global_cache_bias.detach()))
                sw.global_bias = p["global_bias"].data_ptr()
            if self.dino:
                p["dino_bias"] = f32(padc(self.dino_cache_bias.detach()))
                sw.dino_bias = p["dino_bias"].data_ptr()
        p["text_w"] = bf(self.adapter_union_weight)
        p["colscale_text"] = f32(self.logit_scale_text.float().expand(C_))
        sw.text_w, sw.colscale_text = p["text_w"].data_ptr(), p["colscale_text"].data_ptr()
        if self.scoring_precision == "fp32":
            def split3(t):
                t = t.detach().to(device=dev, dtype=torch.float32)
                hi = t.bfloat16()
                r = t - hi.float()
                mid = r.bfloat16()
                return hi, mid, (r - mid.float()).bfloat16()

            def pack6(t):      # weight side of the 3 x bf16 split: [hi | mid | lo | hi | mid | hi] along K
                hi, mid, lo = split3(t)
                return torch.cat([hi, mid, lo, hi, mid, hi], dim=1).contiguous()

            sw32 = _cabi.ScoreWeightsFp32()
            sw32.num_classes, sw32.cache_rows = C_, N
            for i, X in enumerate(("H", "O", "U")):
                p[f"keys6_{X}"] = pack6(padr(getattr(self, f"gen_adapter_{X}_weight").detach()))
                p[f"label3_t_{X}"] = p[f"label_t_{X}"].repeat(1, 3).contiguous()
                sw32.cache_keys6[i], sw32.label3_t[i] = p[f"keys6_{X}"].data_ptr(), p[f"label3_t_{X}"].data_ptr()
                sw32.bias_term[i], sw32.colscale[i] = p[f"bias_term_{X}"].data_ptr(), p[f"colscale_{X}"].data_ptr()
            p["global_keys6"] = (pack6(padr(self.global_cache.detach().t())) if self.clip_global
                                 else torch.zeros(N, 3072, device=dev, dtype=torch.bfloat16))
            sw32.global_keys6 = p["global_keys6"].data_ptr()
            sw32.global_bias_term, sw32.colscale_global = p["global_bias_term"].data_ptr(), p["colscale_global"].data_ptr()
            if self.dino:
                p["dino_keys6"] = pack6(padr(self.dino_cache.detach().t()))
                sw32.dino_keys6 = p["dino_keys6"].data_ptr()
                sw32.dino_bias_term, sw32.colscale_dino = p["dino_bias_term"].data_ptr(), p["colscale_dino"].data_ptr()
            p["text_w6"] = pack6(self.adapter_union_weight)
            sw32.text_w6, sw32.colscale_text = p["text_w6"].data_ptr(), p["colscale_text"].data_ptr()
            p["fp32"] = sw32
        if self.fold_cache:
            # ((f W^T + b) Y) s / L = f (W^T Y s / L) + (b Y) s / L  for every branch; products in fp32, operands to bf16 once
            fw = _cabi.FoldedWeights()
            fw.num_classes = C_
            bias_total = torch.zeros(C_, device=dev)
            blocks = []
            for X in ("H", "O", "U"):
                W, b, Y = (getattr(self, f"gen_adapter_{X}_weight").float().to(dev), getattr(self, f"gen_adapter_{X}_bias").float().to(dev),
                           getattr(self, f"gen_label_{X}").float().to(dev))
                cs = (getattr(self, f"gen_logit_scale_{X}").float().to(dev) / lens(f"sample_lens_{X}", Y))       # (C)
                blocks.append((Y.t() @ W) * cs[:, None])                                              # (C, 512)
                bias_total += (b @ Y) * cs
            blocks[2] = blocks[2] + self.adapter_union_weight.float().to(dev) * self.logit_scale_text.float().to(dev)
            p["fold_pair_w"] = bf(torch.cat(blocks, dim=1))                                            # (C, 1536)
            fw.pair_w = p["fold_pair_w"].data_ptr()
            Yu_d = Yu.to(dev)
            if self.clip_global:
                cs = self.clip_cache_logit.float().to(dev) / lens("global_sample_len", Yu_d)
                p["fold_global_w"] = bf((Yu_d.t() @ self.global_cache.float().to(dev).t()) * cs[:, None])     # (C, 512)
                bias_total += (self.global_cache_bias.float().to(dev) @ Yu_d) * cs
                fw.global_w = p["fold_global_w"].data_ptr()
            if self.dino:
                cs = self.dino_cache_logit.float().to(dev) / lens("dino_sample_len", Yu_d)
                p["fold_dino_w"] = bf((Yu_d.t() @ self.dino_cache.float().to(dev).t()) * cs[:, None])         # (C, 2048)
                bias_total += (self.dino_cache_bias.float().to(dev) @ Yu_d) * cs
                fw.dino_w = p["fold_dino_w"].data_ptr()
            p["fold_bias_total"] = f32(bias_total)
            fw.bias_total = p["fold_bias_total"].data_ptr()
            p["folded"] = fw
        # prior MLP, transposed to (in,out)
        for i, lyr in enumerate(self.priors_downproj.layers):
            p[f"prior_w{i}t"] = f32(lyr.weight.t())
            p[f"prior_b{i}"] = f32(lyr.bias)
        p["object_embedding"] = f32(self.object_embedding)
        # object class -> target classes bitmask (80 x words)
        import numpy as np
        words = (C_ + 31) // 32
        bits = np.zeros((len(self.object_class_to_target_class), words), dtype=np.uint32)
        for o, tars in enumerate(self.object_class_to_target_class):
            for t in tars:
                bits[o, t // 32] |= np.uint32(1 << (t % 32))
        p["table_bits"] = torch.from_numpy(bits.view(np.int32)).to(dev)
        p["table_words"] = words
        p["max_row_len"] = max((len(set(t)) for t in self.object_class_to_target_class), default=0)
        self._packed = (p, sw)
        return self._packed

    # ------------------------------------------------------------------------------------------------------
    # proposals (torch, as in the reference: U:1361-1406) — upstream of the accelerated path
    # ------------------------------------------------------------------------------------------------------
    def prepare_region_proposals(self, results: Sequence[dict]) -> List[dict]:
        from torchvision.ops.boxes import batched_nms
        region_props = []
        for res in results:
            sc, lb, bx = res["scores"], res["labels"], res["boxes"]
            keep = batched_nms(bx, sc, lb, 0.5)
            sc, lb, bx = sc[keep].view(-1), lb[keep].view(-1), bx[keep].view(-1, 4)
            keep = torch.nonzero(sc >= self.box_score_thresh).squeeze(1)
            is_human = lb == self.human_idx
            hum = torch.nonzero(is_human).squeeze(1)
            obj = torch.nonzero(is_human == 0).squeeze(1)
            n_human = int(is_human[keep].sum())
            n_object = len(keep) - n_human

            def select(n_kept, pool, kept_mask):
                if n_kept < self.min_instances:
                    return pool[sc[pool].argsort(descending=True)[: self.min_instances]]
                if n_kept > self.max_instances:
                    return pool[sc[pool].argsort(descending=True)[: self.max_instances]]
                return keep[torch.nonzero(kept_mask).squeeze(1)]

            keep_h = select(n_human, hum, is_human[keep])
            keep_o = select(n_object, obj, is_human[keep] == 0)
            k = torch.cat([keep_h, keep_o])
            # n_human is known on the host here for free; it spares the hot path a device sync
            region_props.append(dict(boxes=bx[k], scores=sc[k], labels=lb[k], n_human=int(keep_h.numel())))
        return region_props

    @torch.no_grad()
    def prepare_region_proposals_batched(self, results: Sequence[dict]):
        """`prepare_region_proposals` (U:1361-1406) for the whole batch in ONE kernel (`hoigen_prepare_proposals`): same
        selection, same order, outputs already in the flat form `launch_packed` takes.  Returns None when the batch is
        not a uniform CUDA batch (different candidate counts per image, more than 256 candidates, CPU tensors): the
        caller then uses the per-image torch form above."""
        if not results:
            return None
        q = int(results[0]["scores"].numel())
        dev = results[0]["scores"].device
        if dev.type != "cuda" or q == 0 or q > 256 or any(int(r["scores"].numel()) != q for r in results):
            return None
        _cabi.init(dev)
        B, cap = len(results), 2 * self.max_instances
        sc = torch.stack([r["scores"].float().reshape(-1) for r in results]).contiguous()
        lb = torch.stack([r["labels"].to(torch.int64).reshape(-1) for r in results]).contiguous()
        bx = torch.stack([r["boxes"].float().reshape(-1, 4) for r in results]).contiguous()
        o_bx = torch.empty(B, cap, 4, device=dev)
        o_sc = torch.empty(B, cap, device=dev)
        o_lb = torch.empty(B, cap, dtype=torch.int64, device=dev)
        counts = torch.empty(B, 2, dtype=torch.int32, device=dev)
        _cabi.call("hoigen_prepare_proposals", sc.data_ptr(), lb.data_ptr(), bx.data_ptr(), B, q, int(self.human_idx),
                   float(self.box_score_thresh), int(self.min_instances), int(self.max_instances), 0.5, o_bx.data_ptr(),
                   o_sc.data_ptr(), o_lb.data_ptr(), counts.data_ptr())
        cnt = counts.cpu()                                   # the layout of the forward is computed on the host
        nh_list = cnt[:, 0].tolist()
        n_list = (cnt[:, 0] + cnt[:, 1]).tolist()
        keep = (torch.arange(cap, device=dev)[None, :] < (counts[:, 0] + counts[:, 1])[:, None].long())
        return o_bx[keep], o_sc[keep], o_lb[keep], n_list, nh_list

    def recover_boxes(self, boxes: torch.Tensor, size: torch.Tensor) -> torch.Tensor:
        """cxcywh in [0,1] -> xyxy pixels (U:1269-1274)."""
        cx, cy, w, h = boxes.unbind(-1)
        b = torch.stack([cx - 0.5 * w, cy - 0.5 * h, cx + 0.5 * w, cy + 0.5 * h], dim=-1)
        hh, ww = size
        return b * torch.stack([ww, hh, ww, hh])

    # ------------------------------------------------------------------------------------------------------
    # the accelerated path: region proposals -> detections
    # ------------------------------------------------------------------------------------------------------
    _PINNED_RING = 8

    def _pinned_slot(self, numel: int):
        """Next slot of the pinned int32 staging ring (layout upload + triplet-offset download of one forward).  A slot is
        reused _PINNED_RING forwards later, after its event (the end of the forward that used it) has completed."""
        ring = self._ws.setdefault("_pinned_ring", [])
        idx = self._ws.get("_pinned_next", 0)
        self._ws["_pinned_next"] = (idx + 1) % self._PINNED_RING
        if len(ring) <= idx:
            ring.append(_PinnedSlot())
        slot = ring[idx]
        if slot.event is not None:
            slot.event.synchronize()
            slot.event = None
        if slot.buf is None or slot.buf.numel() < numel:
            slot.buf = torch.empty(max(numel, 1024), dtype=torch.int32, pin_memory=True)
        slot.generation += 1
        return slot

    def _buf(self, name: str, numel: int, dtype: torch.dtype, device) -> torch.Tensor:
        # per CUDA stream: forwards launched on different streams may overlap on the GPU and must not share scratch
        key = (name, torch.cuda.current_stream(device).cuda_stream) if device.type == "cuda" else name
        t = self._ws.get(key)
        if t is None or t.numel() < numel or t.dtype != dtype or t.device != device:
            t = torch.empty(max(numel, 1), device=device, dtype=dtype)
            self._ws[key] = t
        return t

    @torch.no_grad()
    def forward_from_proposals(self, images_clip: torch.Tensor, region_props: Sequence[dict],
                               dino_image_features: Optional[torch.Tensor] = None, *,
                               return_intermediates: bool = False, encoder_tokens: Optional[torch.Tensor] = None):
        """(B,3,224,224) CLIP images + region proposals (+ L2-normalised DINO features) -> List[dict] as U:1421-1425.

        Equivalent to U:1609-1663 with `prepare_region_proposals` outputs as input (humans lead each image)."""
        return self.finish(self.launch_from_proposals(images_clip, region_props, dino_image_features,
                                                      return_intermediates=return_intermediates,
                                                      encoder_tokens=encoder_tokens))

    @torch.no_grad()
    def launch_from_proposals(self, images_clip: torch.Tensor, region_props: Sequence[dict],
                              dino_image_features: Optional[torch.Tensor] = None, *,
                              return_intermediates: bool = False,
                              encoder_tokens: Optional[torch.Tensor] = None) -> Optional["_PendingForward"]:
        """Enqueue the whole forward on the current stream WITHOUT waiting for it; `finish(handle)` does the path's one
        device->host read (per-image triplet offsets) and builds the detections.  A serving loop calls
        launch(batch i+1) before finish(batch i) so the host never leaves the GPU idle; `forward_from_proposals` is
        launch + finish.  Returns None when no image has a valid pair (U:1660-1662)."""
        n_list = [int(rp["boxes"].shape[0]) for rp in region_props]
        boxes = torch.cat([rp["boxes"] for rp in region_props])
        scores = torch.cat([rp["scores"] for rp in region_props])
        labels = torch.cat([rp["labels"] for rp in region_props])
        if all("n_human" in rp for rp in region_props):
            # produced by this module's own prepare_region_proposals: humans lead by construction (cat([keep_h, keep_o]))
            nh_list = [int(rp["n_human"]) for rp in region_props]
        else:
            # externally supplied proposals: ONE batched device->host read instead of the reference's per-image syncs
            # (U:985-998); it also lets the host validate the labels and apply the reference's humans-first permutation
            lab_h = labels.to(torch.int64).cpu()
            n_obj_classes = min(len(self.object_class_to_target_class), int(self.object_embedding.shape[0]))
            if lab_h.numel() and (int(lab_h.min()) < 0 or int(lab_h.max()) >= n_obj_classes):
                raise IndexError(f"object labels must lie in [0, {n_obj_classes}) (object_class_to_target_class / "
                                 f"object_embedding rows); got [{int(lab_h.min())}, {int(lab_h.max())}]")
            nh_list, perm, off, permuted = [], [], 0, False
            for n in n_list:
                is_h = lab_h[off: off + n] == self.human_idx
                nh = int(is_h.sum())
                nh_list.append(nh)
                if bool(is_h[:nh].all()):
                    perm.append(torch.arange(off, off + n))
                else:   # U:989-996: stable [humans..., others...]; `pairing` then indexes this permuted order while the
                    permuted = True          # returned `boxes` stay as given — exactly what the reference returns
                    perm.append(off + torch.cat([torch.nonzero(is_h).squeeze(1), torch.nonzero(~is_h).squeeze(1)]))
                off += n
            if permuted:
                idx = torch.cat(perm).to(boxes.device)
                boxes, scores, labels = boxes[idx], scores[idx], labels[idx]
        return self.launch_packed(images_clip, boxes, scores, labels, n_list, nh_list, dino_image_features,
                                  return_intermediates=return_intermediates,
                                  image_boxes=[rp["boxes"] for rp in region_props], encoder_tokens=encoder_tokens)

    @torch.no_grad()
    def launch_packed(self, images_clip: torch.Tensor, boxes: torch.Tensor, scores: torch.Tensor, labels: torch.Tensor,
                      n_list: Sequence[int], nh_list: Sequence[int], dino_image_features: Optional[torch.Tensor] = None, *,
                      return_intermediates: bool = False, image_boxes: Optional[Sequence[torch.Tensor]] = None,
                      encoder_tokens: Optional[torch.Tensor] = None):
        """`launch_from_proposals` for a caller that already holds the batch's proposals as flat arrays: boxes (sum n, 4),
        scores (sum n,), labels (sum n,) with image b owning n_list[b] consecutive rows, its nh_list[b] humans first.
        `encoder_tokens` (B*197, 512) fp32, when given, replaces the encoder's output (stage-level parity tests: RoI +
        scoring on IDENTICAL features, north_star's fp32 <= 1e-4 gate)."""
        dev = images_clip.device
        _cabi.init(dev)
        if self._packed is None:
            self.pack_weights()
        p, sw = self._packed
        B = len(n_list)
        n_list, nh_list = [int(v) for v in n_list], [int(v) for v in nh_list]
        Cn = self.num_classes
        img_h, img_w = int(images_clip.shape[-2]), int(images_clip.shape[-1])
        # ---- layout (host): CSR offsets of boxes and pairs -------------------------------------------------
        n_max = max(n_list)
        if n_max > MAX_PRIOR_TOKENS:
            raise ValueError(f"at most {MAX_PRIOR_TOKENS} boxes per image are supported (got {n_max}); the reference caps "
                             f"at 2*max_instances = {2 * self.max_instances}")
        k_list = [(nh * (n - 1) if (nh > 0 and n > 1) else 0) for n, nh in zip(n_list, nh_list)]
        box_off = [0]
        pair_off = [0]
        for n, k in zip(n_list, k_list):
            box_off.append(box_off[-1] + n)
            pair_off.append(pair_off[-1] + k)
        ntot, ktot = box_off[-1], pair_off[-1]
        if ktot == 0:
            return None  # U:1660-1662: no image produced logits (checked before any staging slot is taken or kernel enqueued)
        # pinned staging comes from a small per-module ring: allocating pinned memory per call (cudaHostAlloc) would
        # synchronise the device and serialise launch-ahead callers
        stage = self._pinned_slot(4 * B + 3)
        layout_host = torch.tensor(box_off + pair_off + nh_list, dtype=torch.int32)
        layout = torch.empty(3 * B + 2, dtype=torch.int32, device=dev)
        # carried in kernel parameters: neither a copy-engine memcpy nor a PCIe read that could queue behind a bulk upload
        _cabi.call("hoigen_set_words", layout.data_ptr(), layout_host.data_ptr(), 3 * B + 2)
        d_box_off, d_pair_off, d_nh = layout[: B + 1], layout[B + 1: 2 * B + 2], layout[2 * B + 2:]
        boxes, scores, labels = boxes.float().contiguous(), scores.float().contiguous(), labels.to(torch.int64).contiguous()
        if boxes.shape[0] != ntot or scores.numel() != ntot or labels.numel() != ntot:
            raise ValueError(f"packed proposals hold {boxes.shape[0]} boxes, n_list sums to {ntot}")

        # ---- a3: prior tokens --------------------------------------------------------------------------------
        prior = self._buf("prior", B * n_max * 64, torch.float32, dev)[: B * n_max * 64].view(B, n_max, 64)
        mask = self._buf("mask", B * n_max, torch.uint8, dev)[: B * n_max].view(B, n_max)
        _cabi.call("hoigen_prior_tokens", boxes.data_ptr(), scores.data_ptr(), labels.data_ptr(), d_box_off.data_ptr(),
                   p["object_embedding"].data_ptr(), p["prior_w0t"].data_ptr(), p["prior_b0"].data_ptr(),
                   p["prior_w1t"].data_ptr(), p["prior_b1"].data_ptr(), p["prior_w2t"].data_ptr(), p["prior_b2"].data_ptr(),
                   float(img_w), float(img_h), B, n_max, int(p["object_embedding"].shape[0]), prior.data_ptr(), mask.data_ptr())
        # ---- a4-a7: encoder ------------------------------------------------------------------------------------
        if encoder_tokens is None:
            tokens = self.clip_head.image_encoder.encode_tokens(images_clip, prior, mask)
        else:
            tokens = encoder_tokens.to(device=dev, dtype=torch.float32).contiguous().view(B * TOKENS, 512)
        # ---- a8: DINO features are an input of this path (stock module, as in the reference U:1616-1618) ---------
        if self.dino and dino_image_features is None:
            if self.dino_model is None:
                raise ValueError("dino=True needs `dino_image_features` or a `dino_model`")
            if getattr(self, "_fast_dino", None) is not None:        # opt-in (accelerate_dino): same module, bf16 / graph execution
                dino_image_features = self._fast_dino(images_clip)
            else:
                dino_image_features = self.dino_model(images_clip)
                dino_image_features = dino_image_features / dino_image_features.norm(dim=-1, keepdim=True)
        dino_ptr = dino_image_features.float().contiguous() if self.dino else None

        # ---- a9: RoIAlign + pair assembly --------------------------------------------------------------------------
        single = self._buf("single", ntot * 512, torch.float32, dev)
        union = self._buf("union", ktot * 512, torch.float32, dev)
        pf_bf16 = self._buf("pair_bf16", 3 * ktot * 512, torch.bfloat16, dev)
        fp32_mode = self.scoring_precision == "fp32"
        pf_f32 = self._buf("pair_f32", 3 * ktot * 512, torch.float32, dev) if (return_intermediates or fp32_mode) else None
        spatial_scale = 1.0 / (img_h / 14.0)                                          # U:1027
        roi_w = self._buf("roi_weights", (ntot + ktot) * 32, torch.float32, dev)
        _cabi.call("hoigen_roi_pair_features", tokens.data_ptr(), boxes.data_ptr(), d_box_off.data_ptr(),
                   d_pair_off.data_ptr(), B, ntot, ktot, float(spatial_scale), roi_w.data_ptr(), single.data_ptr(),
                   union.data_ptr(), pf_bf16.data_ptr(), pf_f32.data_ptr() if pf_f32 is not None else None)
        # ---- a10: cache + text logits --------------------------------------------------------------------------------
        N = sw.cache_rows
        ldl = (Cn + 3) // 4 * 4           # padded row pitch: the accumulating GEMM epilogues stay on their float4 path
        logits = self._buf("logits", ktot * ldl, torch.float32, dev)
        if fp32_mode:
            sb32 = _cabi.ScoreBuffersFp32()
            sb32.feat6 = self._buf("feat6", 3 * ktot * 3072, torch.bfloat16, dev).data_ptr()
            sb32.phi = self._buf("phi_f32", ktot * N, torch.float32, dev).data_ptr()
            sb32.phi3 = self._buf("phi3", ktot * 3 * N, torch.bfloat16, dev).data_ptr()
            sb32.phi_img = self._buf("phi_img_f32", B * N, torch.float32, dev).data_ptr()
            sb32.phi_img3 = self._buf("phi_img3", B * 3 * N, torch.bfloat16, dev).data_ptr()
            sb32.g6 = self._buf("g6", B * 3072, torch.bfloat16, dev).data_ptr()
            sb32.d6 = self._buf("d6", B * 6 * 2048, torch.bfloat16, dev).data_ptr()
            sb32.img_logits = self._buf("img_logits", B * Cn, torch.float32, dev).data_ptr()
            sb32.logits, sb32.ld_logits = logits.data_ptr(), ldl
            _cabi.call("hoigen_score_pairs_fp32", C.byref(p["fp32"]), C.byref(sb32), tokens.data_ptr(),
                       dino_ptr.data_ptr() if dino_ptr is not None else None, pf_f32.data_ptr(), d_pair_off.data_ptr(), B, ktot)
        else:
            sb = _cabi.ScoreBuffers()
            sb.pair_feat_bf16 = pf_bf16.data_ptr()
            if Cn <= 128 and not self.fold_cache and os.environ.get("HOIGEN_CACHE_UNFUSED") is None:
                # the three cache branches as ONE fused GEMM-f-GEMM kernel: no (Ktot x N) `phi` buffer at all
                nbytes = int(_cabi.load().hoigen_cache_fused_workspace_bytes(ktot, Cn))
                sb.cache_parts = self._buf("cache_parts", nbytes // 4, torch.float32, dev).data_ptr()
            else:
                sb.phi = self._buf("phi", ktot * N, torch.bfloat16, dev).data_ptr()
            sb.phi_img = self._buf("phi_img", B * N, torch.bfloat16, dev).data_ptr()
            sb.g_bf16 = self._buf("g_bf16", B * 512, torch.bfloat16, dev).data_ptr()
            sb.d_bf16 = self._buf("d_bf16", B * 2048, torch.bfloat16, dev).data_ptr()
            sb.img_logits = self._buf("img_logits", B * Cn, torch.float32, dev).data_ptr()
            sb.logits = logits.data_ptr()
            sb.ld_logits = ldl
            if self.fold_cache:
                _cabi.call("hoigen_score_pairs_folded", C.byref(p["folded"]), C.byref(sb), tokens.data_ptr(),
                           dino_ptr.data_ptr() if dino_ptr is not None else None, d_pair_off.data_ptr(), B, ktot)
            else:
                _cabi.call("hoigen_score_pairs", C.byref(sw), C.byref(sb), tokens.data_ptr(),
                           dino_ptr.data_ptr() if dino_ptr is not None else None, d_pair_off.data_ptr(), B, ktot)
        # ---- a11-a12: prior scores + ordered triplet emission ----------------------------------------------------------
        cap = ktot * p["max_row_len"]
        # the caller owns the outputs (fresh memory per forward, as in the reference) — ONE allocation carved into the five
        # typed arrays: int64 fields first, so every view stays 8-byte aligned
        cap1 = max(cap, 1)
        out_raw = torch.empty(4 * cap1 + (cap1 + 1) // 2 + (B + 2) // 2, device=dev, dtype=torch.int64)
        out_labels, out_objects, out_pairing = out_raw[:cap1], out_raw[cap1: 2 * cap1], out_raw[2 * cap1: 4 * cap1]
        out_scores = out_raw[4 * cap1: 4 * cap1 + (cap1 + 1) // 2].view(torch.float32)[:cap1]
        img_off = out_raw[4 * cap1 + (cap1 + 1) // 2:].view(torch.int32)[: B + 1]
        _cabi.call("hoigen_emit_triplets", logits.data_ptr(), Cn, ldl, scores.data_ptr(), labels.data_ptr(), d_box_off.data_ptr(),
                   d_pair_off.data_ptr(), B, ktot, p["table_bits"].data_ptr(), p["table_words"], int(p["table_bits"].shape[0]),
                   float(self.hyper_lambda),
                   self._buf("emit_counts", ktot, torch.int32, dev).data_ptr(),
                   self._buf("emit_offsets", ktot + 1, torch.int32, dev).data_ptr(),
                   self._buf("emit_pr", ktot, torch.float32, dev).data_ptr(), cap, out_scores.data_ptr(),
                   out_labels.data_ptr(), out_objects.data_ptr(), out_pairing.data_ptr(), img_off.data_ptr())
        # ---- the single device->host read of the path (per-image triplet offsets), asynchronous until finish() --------
        host_off = stage.buf[3 * B + 2: 4 * B + 3]
        _cabi.call("hoigen_copy_words", host_off.data_ptr(), img_off.data_ptr(), B + 1)
        done = torch.cuda.Event()
        done.record()
        stage.event = done
        pend = _PendingForward()
        if image_boxes is None:
            image_boxes = boxes.split(n_list)
        pend.__dict__.update(B=B, img_h=img_h, img_w=img_w, dev=dev, image_boxes=image_boxes, boxes=boxes,
                             box_off=box_off, pair_off=pair_off, ktot=ktot, host_off=host_off, done=done, img_off=img_off,
                             d_box_off=d_box_off,
                             stage=stage, generation=stage.generation, ldl=ldl,
                             out=(out_scores, out_labels, out_objects, out_pairing), prior=prior, mask=mask, tokens=tokens,
                             logits=logits, pf_f32=pf_f32, return_intermediates=return_intermediates)
        return pend

    def finish(self, pend: Optional["_PendingForward"]):
        """Wait for a launched forward and return its detections (List[dict] as U:1421-1425, `.packed` = CSR views)."""
        if pend is None:
            return None
        B, img_h, img_w, dev = pend.B, pend.img_h, pend.img_w, pend.dev
        image_boxes, boxes, box_off, pair_off, ktot = pend.image_boxes, pend.boxes, pend.box_off, pend.pair_off, pend.ktot
        out_scores, out_labels, out_objects, out_pairing = pend.out
        Cn = self.num_classes
        pend.done.synchronize()
        if pend.stage.generation != pend.generation:
            raise RuntimeError(f"more than {self._PINNED_RING - 1} forwards were launched before this one was finished")
        offs = pend.host_off.tolist()
        mtot = offs[-1]
        sizes_m = [offs[b + 1] - offs[b] for b in range(B)]
        size_t = torch.tensor([img_h, img_w], device=dev, dtype=torch.int64)
        # bulk views (one split call per field) instead of 5 Python slices per image
        sc_v = out_scores[:mtot].split(sizes_m)
        lb_v = out_labels[:mtot].split(sizes_m)
        ob_v = out_objects[:mtot].split(sizes_m)
        pr_v = out_pairing[: 2 * mtot].split([2 * m for m in sizes_m])
        detections = DetectionList(
            dict(boxes=image_boxes[b], pairing=pr_v[b].view(2, sizes_m[b]), scores=sc_v[b], labels=lb_v[b],
                 objects=ob_v[b], size=size_t) for b in range(B))
        detections.packed = PackedDetections(scores=out_scores[:mtot], labels=out_labels[:mtot], objects=out_objects[:mtot],
                                             pairing=out_pairing[: 2 * mtot], boxes=boxes, triplet_off=offs, box_off=box_off,
                                             size=(img_h, img_w))
        detections.packed.done = pend.done     # recorded after the last kernel of this forward (for copies on other streams)
        # the same offsets on the device (int32 B+1 each): consumers that stay on the GPU (the wire-format gather) read these
        detections.packed.img_off_dev, detections.packed.box_off_dev = pend.img_off, pend.d_box_off
        if pend.return_intermediates:
            inter = dict(prior=pend.prior, mask=pend.mask.bool(), tokens=pend.tokens.view(B, TOKENS, 512),
                         logits=[pend.logits[: ktot * pend.ldl].view(ktot, pend.ldl)[pair_off[b]: pair_off[b + 1], :Cn]
                                 for b in range(B)],
                         pair_feats=pend.pf_f32[: 3 * ktot * 512].view(3, ktot, 512), pair_off=pair_off, box_off=box_off)
            return detections, inter
        return detections

    # ------------------------------------------------------------------------------------------------------
    def forward(self, images: List, targets: Optional[List[dict]] = None):
        """U:1543-1664, eval branch. images: List[(img_detr (3,H,W), img_clip (3,224,224))]."""
        if self.training:
            raise NotImplementedError("hoigen_b200 accelerates the eval forward; train with the reference module")
        images_orig = [im[0].float() for im in images]
        images_clip = [im[1] for im in images]
        device = images_clip[0].device
        image_sizes = torch.as_tensor([im.size()[-2:] for im in images_clip], device=device)
        if self.detector is None or self.postprocessor is None:
            raise ValueError("UPT.forward needs the injected `detector` and `postprocessor` (U:303-304); "
                             "use forward_from_proposals for precomputed region proposals")
        nested = nested_tensor_from_tensor_list(images_orig)
        if getattr(self, "_fast_detr", None) is not None:          # opt-in (accelerate_detr): the same detector on the repo's kernels
            logits, coords = self._fast_detr(nested.tensors, nested.mask)
            outputs_class, outputs_coord = logits[None], coords[None]
        else:
            features, pos = self.detector.backbone(nested)
            src, mask = features[-1].decompose()
            hs, _ = self.detector.transformer(self.detector.input_proj(src), mask, self.detector.query_embed.weight, pos[-1])
            outputs_class = self.detector.class_embed(hs)
            outputs_coord = self.detector.bbox_embed(hs).sigmoid()
        if self.dataset == "vcoco" and outputs_class.shape[-1] == 92:                 # U:1600-1602
            outputs_class = outputs_class[..., self.reserve_indices.to(outputs_class.device)]
            assert outputs_class.shape[-1] == 81, "reserved shape NOT match 81"
        results = self.postprocessor({"pred_logits": outputs_class[-1], "pred_boxes": outputs_coord[-1]}, image_sizes)
        clip = nested_tensor_from_tensor_list(images_clip).tensors
        batched = self.prepare_region_proposals_batched(results)
        if batched is not None:          # one kernel + one (B,2)-int read-back instead of ~25 torch launches per image
            boxes, scores, labels, n_list, nh_list = batched
            return self.finish(self.launch_packed(clip, boxes, scores, labels, n_list, nh_list))
        region_props = self.prepare_region_proposals(results)
        return self.forward_from_proposals(clip, region_props)
