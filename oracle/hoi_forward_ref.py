"""ORACLE — CPU restatement of HOIGen's eval-mode HOI scoring forward.  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this module;
the product path (hoigen_b200/) never does.  It is a fp32 torch-CPU / numpy restatement (no code copied) of

    U = upt_tip_cache_model_free_finetune_distill3.py      C = CLIP_models_adapter_prior2.py

with each function citing the lines it follows.  Third-party arithmetic on the path that is NOT under
/root/reference: torchvision 0.26.0 `ops.roi_align` (call sites U:1028-1029) — restated here from its published
CPU/CUDA algorithm (roi_align_kernel: aligned=True, adaptive sampling grid, the `< -1 / > size` zeroing rule) —
and torch 2.11 `F.multi_head_attention_forward` / `F.layer_norm` (restated with plain matmul/softmax).

Pinning: oracle/make_golden.py imports the UNMODIFIED reference from /root/reference (builder container only),
runs it on the seeded synthetic state of hoigen_b200/synthetic.py and (a) asserts this restatement matches it
stage by stage, (b) writes the reference's outputs to tests/golden/*.npz.  tests/test_oracle_golden.py re-checks
the restatement against those committed vectors anywhere (no /root/reference needed).
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

ENC = "clip_head.image_encoder."
WIDTH, HEADS, LAYERS, GRID, PATCH = 768, 12, 12, 14, 16


# --------------------------------------------------------------------------------------------------------------
# prior tokens — U:1445-1495 (prior_type 'cbe', prior_method 0) + MLP U:40-52
# --------------------------------------------------------------------------------------------------------------
def prior_tokens(region_props: Sequence[dict], image_hw: Tuple[int, int], head_tensors: Dict[str, torch.Tensor],
                 object_embedding: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    B = len(region_props)
    n_max = max(int(p["boxes"].shape[0]) for p in region_props)
    H, W = image_hw
    raw = torch.zeros(B, n_max, 5 + object_embedding.shape[1])
    mask = torch.ones(B, n_max, dtype=torch.bool)  # True = padding (U:1448, U:1469)
    scale = torch.tensor([W, H, W, H], dtype=torch.float32)
    for b, p in enumerate(region_props):
        n = p["boxes"].shape[0]
        raw[b, :n, 0] = p["scores"]
        raw[b, :n, 1:5] = p["boxes"] / scale
        raw[b, :n, 5:] = object_embedding[p["labels"]]
        mask[b, :n] = False
    x = raw
    for i in range(3):  # 517 -> 128 -> 128 -> 64, ReLU between (U:40-52); padding rows become bias-propagated constants
        x = F.linear(x, head_tensors[f"priors_downproj.layers.{i}.weight"], head_tensors[f"priors_downproj.layers.{i}.bias"])
        if i < 2:
            x = F.relu(x)
    return x, mask


# --------------------------------------------------------------------------------------------------------------
# encoder — C:489-506, blocks C:447-459, adapter C:183-203 + C:51-72
# --------------------------------------------------------------------------------------------------------------
def _ln(x, w, b):
    # C:409-415: computed in fp32, eps = nn.LayerNorm default 1e-5
    mu = x.mean(-1, keepdim=True)
    var = ((x - mu) ** 2).mean(-1, keepdim=True)
    return (x - mu) / torch.sqrt(var + 1e-5) * w + b


def _mha(q_in, k_in, v_in, in_w, in_b, out_w, out_b, heads, key_padding_mask=None):
    """nn.MultiheadAttention forward (eval, batch-major restatement). q_in (B,Lq,E), k_in/v_in (B,Lk,E)."""
    B, Lq, E = q_in.shape
    Lk = k_in.shape[1]
    dh = E // heads
    q = F.linear(q_in, in_w[:E], in_b[:E]).view(B, Lq, heads, dh).transpose(1, 2)
    k = F.linear(k_in, in_w[E:2 * E], in_b[E:2 * E]).view(B, Lk, heads, dh).transpose(1, 2)
    v = F.linear(v_in, in_w[2 * E:], in_b[2 * E:]).view(B, Lk, heads, dh).transpose(1, 2)
    s = (q @ k.transpose(-1, -2)) / math.sqrt(dh)
    if key_padding_mask is not None:
        s = s.masked_fill(key_padding_mask[:, None, None, :], float("-inf"))
    p = torch.softmax(s, dim=-1)
    o = (p @ v).transpose(1, 2).reshape(B, Lq, E)
    return F.linear(o, out_w, out_b)


def adapter_forward(x, prior, mask, sd, blk):
    """C:183-203 with a prior; decoder layer = forward_post C:51-72 (norm1 / self-attn unused)."""
    a = blk + "adaptermlp."
    m = a + "mhsa_layers.0."
    down = F.relu(F.linear(x, sd[a + "down_proj.weight"], sd[a + "down_proj.bias"]))          # C:184-185
    t2 = _mha(down, prior, prior, sd[m + "multihead_attn.in_proj_weight"], sd[m + "multihead_attn.in_proj_bias"],
              sd[m + "multihead_attn.out_proj.weight"], sd[m + "multihead_attn.out_proj.bias"], 2, mask)  # C:63-66
    t = _ln(down + t2, sd[m + "norm2.weight"], sd[m + "norm2.bias"])                            # C:67-68
    t2 = F.linear(F.relu(F.linear(t, sd[m + "linear1.weight"], sd[m + "linear1.bias"])),
                  sd[m + "linear2.weight"], sd[m + "linear2.bias"])                              # C:69
    t = _ln(t + t2, sd[m + "norm3.weight"], sd[m + "norm3.bias"])                               # C:70-71
    up = F.linear(t, sd[a + "up_proj.weight"], sd[a + "up_proj.bias"])                           # C:201
    return up * sd[a + "scale"]                                                                  # C:202


def encoder_forward(images: torch.Tensor, prior: torch.Tensor, mask: torch.Tensor, sd: Dict[str, torch.Tensor],
                    return_layers: bool = False):
    """C:489-506. images (B,3,224,224) -> feat_global (B,512), tokens Y (B,197,512) [feat_local = Y[:,1:] as 14x14 grid]."""
    B = images.shape[0]
    w = sd[ENC + "conv1.weight"]
    # conv1 as im2col GEMM: stride = kernel = 16, no bias (C:476, C:491)
    patches = images.unfold(2, PATCH, PATCH).unfold(3, PATCH, PATCH)           # (B,3,14,14,16,16)
    patches = patches.permute(0, 2, 3, 1, 4, 5).reshape(B, GRID * GRID, 3 * PATCH * PATCH)
    x = patches @ w.reshape(WIDTH, -1).t()                                      # (B,196,768)
    cls = sd[ENC + "class_embedding"].expand(B, 1, WIDTH)
    x = torch.cat([cls, x], dim=1) + sd[ENC + "positional_embedding"]          # C:494-495
    x = _ln(x, sd[ENC + "ln_pre.weight"], sd[ENC + "ln_pre.bias"])              # C:496
    layers = []
    for l in range(LAYERS):
        blk = f"{ENC}transformer.resblocks.{l}."
        x = x + adapter_forward(x, prior, mask, sd, blk)                        # C:454-456
        h = _ln(x, sd[blk + "ln_1.weight"], sd[blk + "ln_1.bias"])
        x = x + _mha(h, h, h, sd[blk + "attn.in_proj_weight"], sd[blk + "attn.in_proj_bias"],
                     sd[blk + "attn.out_proj.weight"], sd[blk + "attn.out_proj.bias"], HEADS)   # C:457
        h = _ln(x, sd[blk + "ln_2.weight"], sd[blk + "ln_2.bias"])
        h = F.linear(h, sd[blk + "mlp.c_fc.weight"], sd[blk + "mlp.c_fc.bias"])
        h = h * torch.sigmoid(1.702 * h)                                        # QuickGELU C:420
        x = x + F.linear(h, sd[blk + "mlp.c_proj.weight"], sd[blk + "mlp.c_proj.bias"])         # C:458
        if return_layers:
            layers.append(x.clone())
    y = _ln(x, sd[ENC + "ln_post.weight"], sd[ENC + "ln_post.bias"]) @ sd[ENC + "proj"]        # C:503-505 (ALL tokens)
    if return_layers:
        return y[:, 0], y, layers
    return y[:, 0], y


def tokens_to_feat_local(y: torch.Tensor) -> torch.Tensor:
    """C:506: (B,197,512) -> (B,512,14,14) view."""
    B = y.shape[0]
    return y[:, 1:].reshape(B, GRID, GRID, -1).permute(0, 3, 1, 2)


# --------------------------------------------------------------------------------------------------------------
# RoIAlign (torchvision.ops.roi_align, aligned=True, sampling_ratio=-1, output 7x7) + mean over bins — U:1027-1037
# --------------------------------------------------------------------------------------------------------------
def roi_align_mean(feat_hw_c: np.ndarray, boxes: np.ndarray, spatial_scale: float, pooled: int = 7) -> np.ndarray:
    """feat_hw_c: (H,W,C) fp32; boxes (R,4) xyxy in image coordinates -> (R,C) = mean over the 7x7 bins of RoIAlign.

    Restates torchvision's roi_align forward (pre_calc bilinear + per-bin average), vectorised over channels,
    accumulating in fp32 in the same (ph, pw, iy, ix) order.
    """
    H, W, C = feat_hw_c.shape
    out = np.zeros((boxes.shape[0], C), dtype=np.float32)
    f32 = np.float32
    for r in range(boxes.shape[0]):
        x1, y1, x2, y2 = [f32(v) * f32(spatial_scale) - f32(0.5) for v in boxes[r]]
        rw, rh = f32(x2 - x1), f32(y2 - y1)                  # aligned=True: no clamp to >= 1
        bw, bh = f32(rw / f32(pooled)), f32(rh / f32(pooled))
        gh = int(math.ceil(float(rh) / pooled))              # adaptive grid (sampling_ratio <= 0)
        gw = int(math.ceil(float(rw) / pooled))
        count = f32(max(gh * gw, 1))
        acc = np.zeros(C, dtype=np.float32)
        for ph in range(pooled):
            for pw in range(pooled):
                bin_sum = np.zeros(C, dtype=np.float32)
                for iy in range(gh):
                    y = f32(y1 + f32(ph) * bh + f32(iy + 0.5) * bh / f32(gh))
                    for ix in range(gw):
                        x = f32(x1 + f32(pw) * bw + f32(ix + 0.5) * bw / f32(gw))
                        if y < -1.0 or y > H or x < -1.0 or x > W:
                            continue
                        yy, xx = max(y, f32(0)), max(x, f32(0))
                        yl, xl = int(yy), int(xx)
                        if yl >= H - 1:
                            yl = yh = H - 1
                            yy = f32(yl)
                        else:
                            yh = yl + 1
                        if xl >= W - 1:
                            xl = xh = W - 1
                            xx = f32(xl)
                        else:
                            xh = xl + 1
                        ly, lx = f32(yy - yl), f32(xx - xl)
                        hy, hx = f32(1) - ly, f32(1) - lx
                        bin_sum += (hy * hx) * feat_hw_c[yl, xl] + (hy * lx) * feat_hw_c[yl, xh] \
                            + (ly * hx) * feat_hw_c[yh, xl] + (ly * lx) * feat_hw_c[yh, xh]
                acc += bin_sum / count
        out[r] = acc / f32(pooled * pooled)                  # .flatten(2).mean(-1)  U:1032-1037 (dropout is identity in eval)
    return out


def pair_indices(n: int, n_h: int) -> Tuple[np.ndarray, np.ndarray]:
    """U:1007-1012: row-major nonzero(x != y and x < n_h) over the n x n meshgrid."""
    xs, ys = [], []
    for x in range(n_h):
        for y in range(n):
            if x != y:
                xs.append(x)
                ys.append(y)
    return np.asarray(xs, dtype=np.int64), np.asarray(ys, dtype=np.int64)


def roi_align_mean_torchvision(feat_hw_c: np.ndarray, boxes: np.ndarray, spatial_scale: float) -> np.ndarray:
    """The reference's own call (U:1028-1037) through the installed torchvision C++ op — used for the CPU-baseline
    timing only (the restatement above is the checker; tests pin the two against each other)."""
    import torchvision
    f = torch.from_numpy(np.ascontiguousarray(feat_hw_c)).permute(2, 0, 1)[None]
    out = torchvision.ops.roi_align(f, [torch.from_numpy(np.ascontiguousarray(boxes))], output_size=(7, 7),
                                    spatial_scale=spatial_scale, aligned=True)
    return out.flatten(2).mean(-1).numpy()


def roi_pair_features(tokens_b: torch.Tensor, props: dict, human_idx: int, image_size: int = 224, roi_impl: str = "restated"):
    """U:981-1057 for one image. tokens_b (197,512). Returns x_keep, y_keep, f_H, f_O, f_U (K,512) (L2-normalised)."""
    roi = roi_align_mean if roi_impl == "restated" else roi_align_mean_torchvision
    boxes = props["boxes"].numpy().astype(np.float32)
    labels = props["labels"].numpy()
    n = boxes.shape[0]
    n_h = int((labels == human_idx).sum())
    assert (labels[:n_h] == human_idx).all(), "humans must lead (prepare_region_proposals U:1398 guarantees it)"
    if n_h == 0 or n <= 1:
        return None
    x_keep, y_keep = pair_indices(n, n_h)
    sub, obj = boxes[x_keep], boxes[y_keep]
    union = np.concatenate([np.minimum(sub[:, :2], obj[:, :2]), np.maximum(sub[:, 2:], obj[:, 2:])], axis=1)  # U:1021-1023
    feat = tokens_b[1:].reshape(GRID, GRID, -1).numpy()
    scale = 1.0 / (image_size / GRID)                                          # U:1027
    uf = torch.from_numpy(roi(feat, union, scale))
    sf = torch.from_numpy(roi(feat, boxes, scale))
    hf, of = sf[x_keep], sf[y_keep]                                            # U:1044-1045
    norm = lambda t: t / t.norm(dim=-1, keepdim=True)                          # U:1048-1050
    return x_keep, y_keep, norm(hf), norm(of), norm(uf)


# --------------------------------------------------------------------------------------------------------------
# scoring — U:1111-1186 (cache_model 'gen_feat', logits_type 'HO+U+T', dino + clip_global)
# --------------------------------------------------------------------------------------------------------------
def scoring_logits(f_h, f_o, f_u, g_b, d_b, T: Dict[str, torch.Tensor], A: Dict[str, torch.Tensor],
                   affinity: str = "linear", beta: float = 10.0, parts: bool = False):
    """g_b (512,) = feat_global[b]/||.||  (U:960);  d_b (2048,) = normalised DINO feature (U:1618).
    affinity='linear' is the reference (phi = f W^T + b, NO exp; U:1156-1158). 'exp' is the textbook Tip-Adapter
    exp(beta * phi) option of north_star, not a parity mode."""
    act = (lambda p: p) if affinity == "linear" else (lambda p: torch.exp(beta * p))
    out = {}
    for X, f in (("H", f_h), ("O", f_o), ("U", f_u)):
        phi = act(f @ T[f"gen_adapter_{X}_weight"].t() + T[f"gen_adapter_{X}_bias"])
        out[X] = (phi @ T[f"gen_label_{X}"]) / A[f"sample_lens_{X}"]          # U:1159-1162
    out["T"] = f_u @ T["adapter_union_weight"].t()                             # U:1163
    K = f_u.shape[0]
    y_u = T["gen_label_U"]                                                     # dino/clip cache values = one_hots_U (U:432,442)
    aff_g = act(g_b @ T["global_cache"] + T["global_cache_bias"])              # U:1133
    out["G"] = ((aff_g @ y_u) / A["global_sample_len"]).expand(K, -1)          # U:1135-1138
    aff_d = act(d_b @ T["dino_cache"] + T["dino_cache_bias"])                  # U:1112
    out["D"] = ((aff_d @ y_u) / A["dino_sample_len"]).expand(K, -1)            # U:1114-1115
    logits = (out["H"] * T["gen_logit_scale_H"] + out["O"] * T["gen_logit_scale_O"] + out["U"] * T["gen_logit_scale_U"]
              + out["T"] * T["logit_scale_text"] + out["G"] * T["clip_cache_logit"] + out["D"] * T["dino_cache_logit"])  # U:1185-1186
    return (logits, out) if parts else logits


# --------------------------------------------------------------------------------------------------------------
# prior scores + triplet emit — U:806-833, U:1408-1427
# --------------------------------------------------------------------------------------------------------------
def prior_scores(x_keep, y_keep, scores: torch.Tensor, labels: torch.Tensor, table: List[List[int]], num_classes: int,
                 hyper_lambda: float) -> torch.Tensor:
    K = len(x_keep)
    prior = torch.zeros(2, K, num_classes)
    s_h = scores[torch.as_tensor(x_keep)].pow(hyper_lambda)                    # eval: p = hyper_lambda (U:814)
    s_o = scores[torch.as_tensor(y_keep)].pow(hyper_lambda)
    for i in range(K):
        for c in table[int(labels[y_keep[i]])]:                                # U:824-831
            prior[0, i, c] = s_h[i]
            prior[1, i, c] = s_o[i]
    return prior


def postprocess(logits: torch.Tensor, prior: torch.Tensor, x_keep, y_keep, labels: torch.Tensor, boxes: torch.Tensor,
                size: Tuple[int, int]) -> dict:
    pr = prior.prod(0)                                                         # U:1417
    x, y = torch.nonzero(pr).unbind(1)                                         # row-major (pair, class)
    xk, yk = torch.as_tensor(x_keep), torch.as_tensor(y_keep)
    return dict(
        boxes=boxes,
        pairing=torch.stack([xk[x], yk[x]]),                                   # U:1422
        scores=torch.sigmoid(logits[x, y]) * pr[x, y],                         # U:1419,1423
        labels=y,
        objects=labels[yk][x],                                                 # U:1424 (objects = labels[y_keep])
        size=torch.as_tensor(size),
    )


# --------------------------------------------------------------------------------------------------------------
# whole path — U:1607-1664 from region proposals onward
# --------------------------------------------------------------------------------------------------------------
def hoi_forward(images: torch.Tensor, region_props: Sequence[dict], dino_feats: torch.Tensor,
                enc_sd: Dict[str, torch.Tensor], head, *, return_intermediates: bool = False,
                affinity: str = "linear", roi_impl: str = "restated"):
    """images (B,3,224,224); region_props as from prepare_region_proposals; dino_feats (B,2048) L2-normalised.
    head: hoigen_b200.synthetic.HeadState-like (tensors, attrs, object_class_to_target_class, num_classes, hyper)."""
    T, A = head.tensors, head.attrs
    hw = (images.shape[-2], images.shape[-1])
    prior, mask = prior_tokens(region_props, hw, T, A["object_embedding"])
    feat_global, tokens = encoder_forward(images, prior, mask, enc_sd)
    g = feat_global / feat_global.norm(dim=-1, keepdim=True)                   # U:960
    dets, inter = [], dict(prior=prior, mask=mask, feat_global=feat_global, tokens=tokens, logits=[], feats=[], priors=[])
    for b, props in enumerate(region_props):
        rp = roi_pair_features(tokens[b], props, head.hyper["human_idx"], hw[0], roi_impl)
        if rp is None:  # U:998-1004: empty detection for this image
            dets.append(dict(boxes=props["boxes"], pairing=torch.zeros(2, 0, dtype=torch.int64), scores=torch.zeros(0),
                             labels=torch.zeros(0, dtype=torch.int64), objects=torch.zeros(0, dtype=torch.int64),
                             size=torch.as_tensor(hw)))
            continue
        x_keep, y_keep, f_h, f_o, f_u = rp
        logits = scoring_logits(f_h, f_o, f_u, g[b], dino_feats[b], T, A, affinity=affinity)
        pri = prior_scores(x_keep, y_keep, props["scores"], props["labels"], head.object_class_to_target_class,
                           head.num_classes, head.hyper["hyper_lambda"])
        dets.append(postprocess(logits, pri, x_keep, y_keep, props["labels"], props["boxes"], hw))
        if return_intermediates:
            inter["logits"].append(logits)
            inter["feats"].append((f_h, f_o, f_u))
            inter["priors"].append(pri)
    return (dets, inter) if return_intermediates else dets


# --------------------------------------------------------------------------------------------------------------
# region proposals — U:1361-1406 (kept in torch on the product side too; restated for the caller-side tests)
# --------------------------------------------------------------------------------------------------------------
def prepare_region_proposals(results: Sequence[dict], human_idx: int, box_score_thresh: float, min_instances: int,
                             max_instances: int) -> List[dict]:
    from torchvision.ops.boxes import batched_nms
    out = []
    for res in results:
        sc, lb, bx = res["scores"], res["labels"], res["boxes"]
        keep = batched_nms(bx, sc, lb, 0.5)
        sc, lb, bx = sc[keep].view(-1), lb[keep].view(-1), bx[keep].view(-1, 4)
        keep = torch.nonzero(sc >= box_score_thresh).squeeze(1)
        is_human = lb == human_idx
        hum = torch.nonzero(is_human).squeeze(1)
        obj = torch.nonzero(is_human == 0).squeeze(1)
        n_human = int(is_human[keep].sum())
        n_object = len(keep) - n_human

        def select(n_kept, pool, kept_mask):
            if n_kept < min_instances:
                return pool[sc[pool].argsort(descending=True)[:min_instances]]
            if n_kept > max_instances:
                return pool[sc[pool].argsort(descending=True)[:max_instances]]
            return keep[torch.nonzero(kept_mask).squeeze(1)]

        keep_h = select(n_human, hum, is_human[keep])
        keep_o = select(n_object, obj, is_human[keep] == 0)
        k = torch.cat([keep_h, keep_o])
        out.append(dict(boxes=bx[k], scores=sc[k], labels=lb[k]))
    return out


# ----------------------------------------------------------------------------------------------------------------------
# f3: the proposal stage restated WITHOUT torchvision — batched_nms' coordinate trick
# (torchvision/ops/boxes.py, 0.26.0: offsets = idxs.to(boxes) * (boxes.max() + 1); nms(boxes + offsets, ...)) and the
# greedy NMS of torchvision/csrc/ops/cpu/nms_kernel.cpp (stable sort by score descending; areas = (x2-x1)*(y2-y1);
# ovr = inter / (iarea + areas[j] - inter) > thr suppresses j), every operation rounded to fp32 separately; then the
# threshold / min / max-instance selection of U:1361-1406.  Checks hoigen_prepare_proposals.
# ----------------------------------------------------------------------------------------------------------------------
def batched_nms_ref(boxes: torch.Tensor, scores: torch.Tensor, idxs: torch.Tensor, iou_threshold: float) -> torch.Tensor:
    if boxes.numel() == 0:
        return torch.empty((0,), dtype=torch.int64)
    b = boxes.float().numpy()
    mx = np.float32(b.max())
    off = idxs.numpy().astype(np.float32) * np.float32(mx + np.float32(1))
    s = (b + off[:, None]).astype(np.float32)
    x1, y1, x2, y2 = s[:, 0], s[:, 1], s[:, 2], s[:, 3]
    areas = ((x2 - x1).astype(np.float32) * (y2 - y1).astype(np.float32)).astype(np.float32)
    order = np.argsort(-scores.float().numpy(), kind="stable")
    n = len(order)
    supp = np.zeros(n, dtype=bool)
    keep = []
    thr = np.float32(iou_threshold)
    for _i in range(n):
        i = order[_i]
        if supp[i]:
            continue
        keep.append(i)
        for _j in range(_i + 1, n):
            j = order[_j]
            if supp[j]:
                continue
            w = np.float32(max(np.float32(0), np.float32(min(x2[i], x2[j]) - max(x1[i], x1[j]))))
            h = np.float32(max(np.float32(0), np.float32(min(y2[i], y2[j]) - max(y1[i], y1[j]))))
            inter = np.float32(w * h)
            ovr = np.float32(inter / np.float32(np.float32(areas[i] + areas[j]) - inter))
            if ovr > thr:
                supp[j] = True
    return torch.tensor(keep, dtype=torch.int64)


def prepare_region_proposals_ref(results: Sequence[dict], human_idx: int, box_score_thresh: float, min_instances: int,
                                 max_instances: int) -> List[dict]:
    """U:1361-1406 with batched_nms_ref: after NMS everything is in descending-score order, so every branch of the
    min / max-instance rule is "the first k humans (objects) in that order", k = min_instances (or all there are) if
    fewer than min_instances pass the score threshold, max_instances if more than max_instances do, else the count."""
    out = []
    for res in results:
        sc, lb, bx = res["scores"], res["labels"], res["boxes"]
        keep = batched_nms_ref(bx, sc, lb, 0.5)
        sc, lb, bx = sc[keep].view(-1), lb[keep].view(-1), bx[keep].view(-1, 4)
        is_h = lb == human_idx
        sel = []
        for mask in (is_h, ~is_h):
            pool = torch.nonzero(mask).squeeze(1)
            n_ok = int((sc[pool] >= box_score_thresh).sum())
            k = min_instances if n_ok < min_instances else (max_instances if n_ok > max_instances else n_ok)
            sel.append(pool[:k])
        k = torch.cat(sel)
        out.append(dict(boxes=bx[k], scores=sc[k], labels=lb[k], n_human=int(sel[0].numel())))
    return out


def synthetic_detr_results(batch: int, seed: int, q: int = 100, tie_every: int = 0) -> List[dict]:
    """DETR-like raw detections for the proposal-stage tests: q candidates per image clustered around a few centres (so
    NMS has work to do inside a class and must NOT suppress across classes), ~35 % humans, scores spread around the
    threshold.  Image b % 4 == 1 has no human at all, b % 4 == 2 has every score below the threshold (min-instance
    branch), b % 4 == 3 has every score high (max-instance branch).  tie_every > 0 repeats scores (stable-order ties)."""
    g = torch.Generator().manual_seed(seed)
    out = []
    for b in range(batch):
        centres = torch.rand(12, 2, generator=g) * 160 + 30
        which = torch.randint(0, 12, (q,), generator=g)
        ctr = centres[which] + torch.randn(q, 2, generator=g) * 6
        wh = torch.rand(q, 2, generator=g) * 50 + 15
        boxes = torch.cat([ctr - wh / 2, ctr + wh / 2], 1)
        labels = torch.tensor([1, 2, 3, 40, 79])[torch.randint(0, 5, (q,), generator=g)]   # few classes: NMS inside each;
        if b % 4 != 1:                                   # large ids: the coordinate trick's fp32 rounding matters
            labels[torch.rand(q, generator=g) < 0.35] = 0
        scores = torch.rand(q, generator=g)
        if b % 4 == 2:
            scores = scores * 0.19
        elif b % 4 == 3:
            scores = 0.5 + scores * 0.5
        if tie_every:
            scores[::tie_every] = scores[0]
        out.append(dict(scores=scores, labels=labels, boxes=boxes))
    return out
