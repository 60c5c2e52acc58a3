"""Seeded synthetic state + inputs for the HOI scoring forward (SURVEY.md §8d "Synthetic inputs").

Everything is generated on the CPU with explicit torch.Generator seeds so that the very same tensors can be
re-created (a) in the builder container where the real reference is importable (oracle/make_golden.py),
(b) on the GPU box for the parity tests and (c) inside bench.py.  No dataset, checkpoint or network access.

Names follow the reference's state_dict: `clip_head.image_encoder.*`,
`priors_downproj.layers.*`, `gen_adapter_{U,H,O}_weight` ... so the same dict loads into the reference UPT
(upt_tip_cache_model_free_finetune_distill3.py:299-604) and into hoigen_b200.detector.UPT.
"""
from __future__ import annotations

import json
import math
from dataclasses import dataclass, field
from pathlib import Path
from typing import Dict, List

import torch

WIDTH, LAYERS, HEADS, PATCH, RES, OUT_DIM = 768, 12, 12, 16, 224, 512
TOKENS = (RES // PATCH) ** 2 + 1  # 197
ADAPTER_DIM = 64
ENC_PREFIX = "clip_head.image_encoder."

_TABLE_PATH = Path(__file__).resolve().parent / "data" / "object_tables.json"


def _gen(seed: int) -> torch.Generator:
    g = torch.Generator(device="cpu")
    g.manual_seed(seed)
    return g


def _randn(g, *shape, std=1.0):
    return torch.randn(*shape, generator=g, dtype=torch.float32) * std


def make_encoder_state(seed: int = 0) -> Dict[str, torch.Tensor]:
    """Random ViT-B/16 + InsAdapter parameters (CLIP_models_adapter_prior2.py:463-506 shapes).

    Unlike the reference's own init (up_proj = 0, scale = 1e-9, zero biases: C:157,172) every tensor the
    forward reads is made non-trivial, otherwise the adapter and bias paths would be untested.
    """
    g = _gen(seed)
    sd: Dict[str, torch.Tensor] = {}
    p = ENC_PREFIX
    s = WIDTH ** -0.5
    sd[p + "class_embedding"] = _randn(g, WIDTH, std=s)
    sd[p + "positional_embedding"] = _randn(g, TOKENS, WIDTH, std=s)
    sd[p + "proj"] = _randn(g, WIDTH, OUT_DIM, std=s)
    sd[p + "conv1.weight"] = _randn(g, WIDTH, 3, PATCH, PATCH, std=(3 * PATCH * PATCH) ** -0.5)
    for ln in ("ln_pre", "ln_post"):
        sd[p + ln + ".weight"] = 1.0 + _randn(g, WIDTH, std=0.1)
        sd[p + ln + ".bias"] = _randn(g, WIDTH, std=0.1)
    attn_std = WIDTH ** -0.5
    proj_std = (WIDTH ** -0.5) * ((2 * LAYERS) ** -0.5)
    fc_std = (2 * WIDTH) ** -0.5
    for i in range(LAYERS):
        b = f"{p}transformer.resblocks.{i}."
        sd[b + "attn.in_proj_weight"] = _randn(g, 3 * WIDTH, WIDTH, std=attn_std)
        sd[b + "attn.in_proj_bias"] = _randn(g, 3 * WIDTH, std=0.02)
        sd[b + "attn.out_proj.weight"] = _randn(g, WIDTH, WIDTH, std=proj_std)
        sd[b + "attn.out_proj.bias"] = _randn(g, WIDTH, std=0.02)
        for ln in ("ln_1", "ln_2"):
            sd[b + ln + ".weight"] = 1.0 + _randn(g, WIDTH, std=0.1)
            sd[b + ln + ".bias"] = _randn(g, WIDTH, std=0.1)
        sd[b + "mlp.c_fc.weight"] = _randn(g, 4 * WIDTH, WIDTH, std=fc_std)
        sd[b + "mlp.c_fc.bias"] = _randn(g, 4 * WIDTH, std=0.02)
        sd[b + "mlp.c_proj.weight"] = _randn(g, WIDTH, 4 * WIDTH, std=proj_std)
        sd[b + "mlp.c_proj.bias"] = _randn(g, WIDTH, std=0.02)
        a = b + "adaptermlp."
        sd[a + "scale"] = 1.0 + _randn(g, WIDTH, std=0.1)
        sd[a + "down_proj.weight"] = _randn(g, ADAPTER_DIM, WIDTH, std=WIDTH ** -0.5)
        sd[a + "down_proj.bias"] = _randn(g, ADAPTER_DIM, std=0.02)
        sd[a + "up_proj.weight"] = _randn(g, WIDTH, ADAPTER_DIM, std=0.02)
        sd[a + "up_proj.bias"] = _randn(g, WIDTH, std=0.02)
        m = a + "mhsa_layers.0."
        sd[m + "multihead_attn.in_proj_weight"] = _randn(g, 3 * ADAPTER_DIM, ADAPTER_DIM, std=ADAPTER_DIM ** -0.5)
        sd[m + "multihead_attn.in_proj_bias"] = _randn(g, 3 * ADAPTER_DIM, std=0.02)
        sd[m + "multihead_attn.out_proj.weight"] = _randn(g, ADAPTER_DIM, ADAPTER_DIM, std=ADAPTER_DIM ** -0.5)
        sd[m + "multihead_attn.out_proj.bias"] = _randn(g, ADAPTER_DIM, std=0.02)
        sd[m + "linear1.weight"] = _randn(g, 2 * ADAPTER_DIM, ADAPTER_DIM, std=ADAPTER_DIM ** -0.5)
        sd[m + "linear1.bias"] = _randn(g, 2 * ADAPTER_DIM, std=0.02)
        sd[m + "linear2.weight"] = _randn(g, ADAPTER_DIM, 2 * ADAPTER_DIM, std=(2 * ADAPTER_DIM) ** -0.5)
        sd[m + "linear2.bias"] = _randn(g, ADAPTER_DIM, std=0.02)
        for ln in ("norm2", "norm3"):
            sd[m + ln + ".weight"] = 1.0 + _randn(g, ADAPTER_DIM, std=0.1)
            sd[m + ln + ".bias"] = _randn(g, ADAPTER_DIM, std=0.1)
    return sd


def load_object_tables() -> dict:
    """Object -> target-class tables derived from the reference's annotation JSONs by
    oracle/make_tables.py (hicodet/instances_test2015.json['correspondence'],
    vcoco/instances_vcoco_test.json); committed as package data (small integer lists)."""
    with open(_TABLE_PATH) as f:
        return json.load(f)


def object_table(num_classes: int) -> List[List[int]]:
    t = load_object_tables()
    key = {117: "hico_object_to_verb", 600: "hico_object_to_interaction", 24: "vcoco_object_to_action"}[num_classes]
    return t[key]


@dataclass
class HeadState:
    """Everything outside the encoder that the eval forward reads."""
    tensors: Dict[str, torch.Tensor]                 # state_dict entries (reference parameter names)
    attrs: Dict[str, torch.Tensor]                   # plain attributes (not in the checkpoint)
    object_class_to_target_class: List[List[int]]
    num_classes: int
    hyper: dict = field(default_factory=dict)


def make_head_state(num_classes: int = 117, cache_rows: int = 4096, seed: int = 2, *, hyper_lambda: float = 2.8,
                    box_score_thresh: float = 0.2, min_instances: int = 3, max_instances: int = 15,
                    human_idx: int = 0) -> HeadState:
    """Cache keys / labels / biases / text classifier / prior MLP:
    keys unit-norm randn, biases -1 + 0.01 randn, Y = block one-hot + 10% random extra ones, s = Y.sum(0)."""
    g = _gen(seed)
    C, N = num_classes, cache_rows
    t: Dict[str, torch.Tensor] = {}
    a: Dict[str, torch.Tensor] = {}

    def unit(x, dim=-1):
        return x / x.norm(dim=dim, keepdim=True)

    def labels():
        y = torch.zeros(N, C)
        rows = torch.arange(N)
        y[rows, (rows * C) // N] = 1.0            # block one-hot: ~N/C consecutive rows per class
        extra = torch.rand(N, generator=g) < 0.10  # ~10% of the rows get one extra positive (multi-hot)
        cols = torch.randint(0, C, (N,), generator=g)
        y[rows[extra], cols[extra]] = 1.0
        return y

    for X in ("U", "H", "O"):
        t[f"gen_adapter_{X}_weight"] = unit(_randn(g, N, OUT_DIM))
        t[f"gen_adapter_{X}_bias"] = -1.0 + _randn(g, N, std=0.01)
        t[f"gen_label_{X}"] = labels()
        t[f"gen_logit_scale_{X}"] = torch.tensor(math.log(1 / 0.07)) + _randn(g, 1, std=0.05)[0]
        a[f"sample_lens_{X}"] = t[f"gen_label_{X}"].sum(0)
    t["adapter_union_weight"] = unit(_randn(g, C, OUT_DIM))
    t["logit_scale_text"] = torch.tensor(math.log(1 / 0.07)) + _randn(g, 1, std=0.05)[0]
    t["dino_cache"] = unit(_randn(g, 2048, N), dim=0)
    t["dino_cache_bias"] = -1.0 + _randn(g, N, std=0.01)
    t["dino_cache_logit"] = torch.tensor(math.log(1 / 0.07)) + _randn(g, 1, std=0.05)[0]
    t["global_cache"] = unit(_randn(g, OUT_DIM, N), dim=0)
    t["global_cache_bias"] = -1.0 + _randn(g, N, std=0.01)
    t["clip_cache_logit"] = torch.tensor(math.log(1 / 0.07)) + _randn(g, 1, std=0.05)[0]
    # U:432,442-445: the dino / global cache VALUES are the union labels
    a["dino_sample_len"] = a["sample_lens_U"].clone()
    a["global_sample_len"] = a["sample_lens_U"].clone()
    # prior MLP 517 -> 128 -> 128 -> 64 (U:520, U:40-52)
    dims = [OUT_DIM + 5, 128, 128, 64]
    for i in range(3):
        t[f"priors_downproj.layers.{i}.weight"] = _randn(g, dims[i + 1], dims[i], std=dims[i] ** -0.5)
        t[f"priors_downproj.layers.{i}.bias"] = _randn(g, dims[i + 1], std=0.02)
    # unnormalised CLIP text features of the 80 object prompts (U:1704-1707); magnitude ~ 10 in real CLIP
    a["object_embedding"] = _randn(g, 80, OUT_DIM, std=0.4)
    return HeadState(
        tensors=t, attrs=a, object_class_to_target_class=object_table(num_classes), num_classes=C,
        hyper=dict(hyper_lambda=hyper_lambda, box_score_thresh=box_score_thresh, min_instances=min_instances,
                   max_instances=max_instances, human_idx=human_idx),
    )


def make_images(batch: int, seed: int = 1) -> torch.Tensor:
    return _randn(_gen(seed), batch, 3, RES, RES)


def make_dino_features(batch: int, seed: int = 5) -> torch.Tensor:
    """Stand-in for the L2-normalised DINO ResNet-50 features (U:1616-1618): row a8 is an INPUT of this path."""
    x = _randn(_gen(seed), batch, 2048)
    return x / x.norm(dim=-1, keepdim=True)


def make_boxes(image_index: int, n_human: int = 8, n_object: int = 8, *, size: int = RES) -> dict:
    """NMS-safe grid boxes (pairwise IoU < 0.5, inside the image), humans first.

    Returns the dict `prepare_region_proposals` would produce: boxes (n,4) xyxy fp32, scores (n,), labels (n,) i64.
    """
    g = _gen(100 + image_index)
    n = n_human + n_object
    side = int(math.ceil(math.sqrt(n)))
    cell = size / side
    lo, hi = (20.0, 26.0) if n <= 16 else (12.0, 17.0)
    idx = torch.randperm(side * side, generator=g)[:n]
    cy = (idx // side).float() * cell + cell / 2
    cx = (idx % side).float() * cell + cell / 2
    hw = lo + (hi - lo) * torch.rand(n, generator=g)
    hh = lo + (hi - lo) * torch.rand(n, generator=g)
    boxes = torch.stack([cx - hw, cy - hh, cx + hw, cy + hh], dim=1).clamp_(0, size)
    labels = torch.cat([torch.zeros(n_human, dtype=torch.int64),
                        torch.randint(1, 80, (n_object,), generator=g)])
    scores = 0.3 + 0.69 * torch.rand(n, generator=g)
    # prepare_region_proposals (U:1366-1398) returns each group in descending-score order (batched_nms sorts)
    order = torch.cat([scores[:n_human].argsort(descending=True), n_human + scores[n_human:].argsort(descending=True)])
    return dict(boxes=boxes[order], scores=scores[order], labels=labels[order])


def make_region_props(batch: int, n_human: int = 8, n_object: int = 8, *, ragged: bool = False, seed: int = 0) -> List[dict]:
    props = []
    for b in range(batch):
        nh, no = n_human, n_object
        if ragged:  # vary the instance counts per image (exercises CSR offsets / n_max padding)
            nh = max(1, n_human - (b % 3))
            no = max(1, n_object - ((2 * b) % 5))
        props.append(make_boxes(seed + b, nh, no))
    return props


# ----------------------------------------------------------------------------------------------------------------------
# BASELINE configs[2] ("UC zero-shot (uc0) with VAE-synthesized unseen-class union features appended to a 16k x 512
# cache"): only the cache CONTENT differs from configs[1] (the eval forward never reads `zs` / `zs_type`).  The appended
# rows follow main_tip_finetune.py:749-824 for the 120 held-out uc0 HOI classes:
#     z ~ N(0,1) (R,512) -> Generator 512->4096->512 (M:247-261) -> bias
#     prompt = [SOS emb | ctx (n_ctx,512) + bias | class-name + EOS + padding embs]  (PromptLearner_*.forward, M:106-115)
#     frozen CLIP text encoder (TextEncoder M:262-279: + positional_embedding, 12 causal blocks, ln_final, EOT row @
#     text_projection) -> L2 norm -> mlp_net 512->512->512->512 (M:313-324)
# with random-init modules (seed 3) and synthetic token embeddings (no BPE vocabulary on the GPU box): offline cache
# construction, torch-CPU, not part of the timed path.  oracle/make_golden.py pins this restatement against the
# reference's own Generator / TextEncoder / mlp_net classes.
# ----------------------------------------------------------------------------------------------------------------------
TEXT_WIDTH, TEXT_LAYERS, TEXT_HEADS, TEXT_CTX = 512, 12, 8, 77


def make_text_tower_state(seed: int = 3) -> Dict[str, torch.Tensor]:
    """Random CLIP text tower (vanilla CLIP names: transformer.resblocks.{i}.*, positional_embedding, ln_final.*,
    text_projection) with CLIP's initialisation scales."""
    g = _gen(seed)
    W, L = TEXT_WIDTH, TEXT_LAYERS
    sd = {"positional_embedding": _randn(g, TEXT_CTX, W, std=0.01), "text_projection": _randn(g, W, OUT_DIM, std=W ** -0.5),
          "ln_final.weight": 1.0 + _randn(g, W, std=0.05), "ln_final.bias": _randn(g, W, std=0.05)}
    proj_std, attn_std, fc_std = (W ** -0.5) * ((2 * L) ** -0.5), W ** -0.5, (2 * W) ** -0.5
    for i in range(L):
        b = f"transformer.resblocks.{i}."
        sd[b + "attn.in_proj_weight"] = _randn(g, 3 * W, W, std=attn_std)
        sd[b + "attn.in_proj_bias"] = _randn(g, 3 * W, std=0.01)
        sd[b + "attn.out_proj.weight"] = _randn(g, W, W, std=proj_std)
        sd[b + "attn.out_proj.bias"] = _randn(g, W, std=0.01)
        for ln in ("ln_1", "ln_2"):
            sd[b + ln + ".weight"] = 1.0 + _randn(g, W, std=0.05)
            sd[b + ln + ".bias"] = _randn(g, W, std=0.05)
        sd[b + "mlp.c_fc.weight"] = _randn(g, 4 * W, W, std=fc_std)
        sd[b + "mlp.c_fc.bias"] = _randn(g, 4 * W, std=0.01)
        sd[b + "mlp.c_proj.weight"] = _randn(g, W, 4 * W, std=proj_std)
        sd[b + "mlp.c_proj.bias"] = _randn(g, W, std=0.01)
    return sd


def make_generator_state(seed: int = 3, n_ctx: int = 5) -> Dict[str, torch.Tensor]:
    """Generator (`net.0/2`: N(0,0.02) weights, zero biases = weights_init M:44-51), prompt context `ctx`, mlp_net
    (`net.0/2/4`, torch Linear-style uniform init)."""
    g = _gen(seed + 1000)
    sd = {"gen.net.0.weight": _randn(g, 4096, 512, std=0.02), "gen.net.0.bias": torch.zeros(4096),
          "gen.net.2.weight": _randn(g, 512, 4096, std=0.02), "gen.net.2.bias": torch.zeros(512),
          "ctx": _randn(g, n_ctx, 512, std=0.02)}
    for j in (0, 2, 4):
        bound = 512 ** -0.5
        sd[f"mlp.net.{j}.weight"] = (torch.rand(512, 512, generator=g) * 2 - 1) * bound
        sd[f"mlp.net.{j}.bias"] = (torch.rand(512, generator=g) * 2 - 1) * bound
    return sd


def text_encoder_forward(prompts: torch.Tensor, eot: torch.Tensor, tw: Dict[str, torch.Tensor]) -> torch.Tensor:
    """TextEncoder.forward (M:270-279) over (R,77,512) prompt embeddings; `eot` (R,) = position of the EOT token."""
    F = torch.nn.functional
    R, T, W = prompts.shape
    hd = W // TEXT_HEADS
    causal = torch.full((T, T), float("-inf")).triu_(1)
    x = prompts + tw["positional_embedding"]
    for i in range(TEXT_LAYERS):
        b = f"transformer.resblocks.{i}."
        h = F.layer_norm(x, (W,), tw[b + "ln_1.weight"], tw[b + "ln_1.bias"])
        q, k, v = F.linear(h, tw[b + "attn.in_proj_weight"], tw[b + "attn.in_proj_bias"]).view(R, T, 3, TEXT_HEADS, hd).unbind(2)
        att = torch.softmax(torch.einsum("rthd,rshd->rhts", q * hd ** -0.5, k) + causal, dim=-1)
        o = torch.einsum("rhts,rshd->rthd", att, v).reshape(R, T, W)
        x = x + F.linear(o, tw[b + "attn.out_proj.weight"], tw[b + "attn.out_proj.bias"])
        h = F.layer_norm(x, (W,), tw[b + "ln_2.weight"], tw[b + "ln_2.bias"])
        h = F.linear(h, tw[b + "mlp.c_fc.weight"], tw[b + "mlp.c_fc.bias"])
        x = x + F.linear(h * torch.sigmoid(1.702 * h), tw[b + "mlp.c_proj.weight"], tw[b + "mlp.c_proj.bias"])
    x = F.layer_norm(x, (W,), tw["ln_final.weight"], tw["ln_final.bias"])
    return x[torch.arange(R), eot] @ tw["text_projection"]


def generated_rows(targets: torch.Tensor, num_names: int, seed: int = 3, n_ctx: int = 5, chunk: int = 64,
                   return_inputs: bool = False):
    """(R,512) generator-synthesised cache rows for class indices `targets` (R,) in [0, num_names)."""
    F = torch.nn.functional
    tw, gs = make_text_tower_state(seed), make_generator_state(seed, n_ctx)
    g = _gen(seed + 2000)
    # synthetic token embeddings of "X X X X X <name>." : SOS, (ctx), 1-4 name tokens, '.', EOT, padding (zeros row of the
    # embedding table in real CLIP is a learnt row too; here every position is a random row, as in a random-init model)
    name_len = 1 + (torch.arange(num_names) % 4)
    table = _randn(g, num_names, TEXT_CTX, TEXT_WIDTH, std=0.02)
    eot_pos = 1 + n_ctx + name_len + 1                      # SOS + ctx + name + '.' -> EOT
    z = _randn(g, targets.numel(), 512)
    out = []
    for s0 in range(0, targets.numel(), chunk):
        t = targets[s0: s0 + chunk]
        bias = F.linear(F.relu(F.linear(z[s0: s0 + chunk], gs["gen.net.0.weight"], gs["gen.net.0.bias"])),
                        gs["gen.net.2.weight"], gs["gen.net.2.bias"])
        emb = table[t]
        prompts = torch.cat([emb[:, :1], gs["ctx"][None] + bias[:, None, :], emb[:, 1 + n_ctx:]], dim=1)
        f = text_encoder_forward(prompts, eot_pos[t], tw)
        f = f / f.norm(dim=-1, keepdim=True)
        for j in (0, 2):
            f = F.relu(F.linear(f, gs[f"mlp.net.{j}.weight"], gs[f"mlp.net.{j}.bias"]))
        out.append(F.linear(f, gs["mlp.net.4.weight"], gs["mlp.net.4.bias"]))
    rows = torch.cat(out)
    if return_inputs:
        return rows, dict(z=z, table=table, eot=eot_pos, text=tw, gen=gs)
    return rows


def make_head_state_uc0(cache_rows: int = 16384, seed: int = 2, gen_seed: int = 3) -> HeadState:
    """configs[2]: the configs[1] head with N rows whose LAST 120 rows per branch are generator-synthesised features of
    the 120 held-out uc0 HOI classes (union rows from the 600 HOI prompts, human rows from the single 'person' prompt
    family, object rows from the 80 object prompts), labelled with each class's verb."""
    head = make_head_state(117, cache_rows, seed)
    tabs = load_object_tables()
    uc0 = tabs["hico_unseen_uc0"]
    corr = {h: (o, v) for h, o, v in tabs["hico_correspondence"]}
    R = len(uc0)
    hoi = torch.tensor(uc0)
    obj = torch.tensor([corr[h][0] for h in uc0])
    verb = torch.tensor([corr[h][1] for h in uc0])
    rows = {"U": generated_rows(hoi, 600, gen_seed, n_ctx=5), "H": generated_rows(obj, 80, gen_seed + 1, n_ctx=4),
            "O": generated_rows(obj, 80, gen_seed + 2, n_ctx=4)}
    y = torch.zeros(R, 117)
    y[torch.arange(R), verb] = 1.0
    for X in ("U", "H", "O"):
        head.tensors[f"gen_adapter_{X}_weight"][-R:] = rows[X]
        head.tensors[f"gen_label_{X}"][-R:] = y
        head.attrs[f"sample_lens_{X}"] = head.tensors[f"gen_label_{X}"].sum(0)
    head.attrs["dino_sample_len"] = head.attrs["sample_lens_U"].clone()
    head.attrs["global_sample_len"] = head.attrs["sample_lens_U"].clone()
    return head
