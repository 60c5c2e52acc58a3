"""Image sharding + the one collective of the path: a variable-length gather of per-image detections.

The reference evaluates on a single GPU (batch 1, sequential sampler: main_tip_finetune.py:383-388) and has no
detection gather; its generic helper is a pickle-based padded all_gather (pocket/utils/distributed.py:103-143).
Here every rank scores a contiguous shard of the images (weights, cache keys and text embeddings replicated) and
the detections are exchanged as flat CSR tensors: counts first, then one padded payload per dtype — NCCL over
NVLink/NVSwitch on the GPU box, gloo in the CPU tests.
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist


def shard_range(total: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous shard [lo, hi) of `total` images for `rank` (sizes differ by at most one)."""
    base, rem = divmod(total, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def pack_detections(dets: Sequence[Optional[dict]]):
    """List[dict] (U:1421-1425) -> (counts (B,2) i64 [triplets, boxes], floats (M,1)+(n,4) packed f32, ints (M,4) i64)."""
    dev = None
    for d in dets:
        if d is not None:
            dev = d["scores"].device
            break
    if dev is None:
        dev = torch.device("cpu")
    counts = torch.tensor([[0 if d is None else d["scores"].numel(), 0 if d is None else d["boxes"].shape[0]] for d in dets],
                          dtype=torch.int64, device=dev).view(-1, 2)
    live = [d for d in dets if d is not None]
    if live:
        scores = torch.cat([d["scores"].float() for d in live])
        boxes = torch.cat([d["boxes"].float().reshape(-1) for d in live])
        ints = torch.cat([torch.stack([d["labels"], d["objects"], d["pairing"][0], d["pairing"][1]], dim=1) for d in live])
    else:
        scores = torch.zeros(0, device=dev)
        boxes = torch.zeros(0, device=dev)
        ints = torch.zeros(0, 4, dtype=torch.int64, device=dev)
    return counts, torch.cat([scores, boxes]), ints.to(torch.int64)


def unpack_detections(counts: torch.Tensor, floats: torch.Tensor, ints: torch.Tensor, size=(224, 224)) -> List[dict]:
    counts = counts.cpu().tolist()
    m_tot = sum(c[0] for c in counts)
    scores, boxes = floats[:m_tot], floats[m_tot:]
    out, mo, bo = [], 0, 0
    for m, n in counts:
        i = ints[mo: mo + m]
        out.append(dict(boxes=boxes[bo: bo + 4 * n].view(n, 4), pairing=torch.stack([i[:, 2], i[:, 3]]) if m else
                        torch.zeros(2, 0, dtype=torch.int64, device=ints.device),
                        scores=scores[mo: mo + m], labels=i[:, 0], objects=i[:, 1],
                        size=torch.tensor(size, dtype=torch.int64, device=ints.device)))
        mo += m
        bo += 4 * n
    return out


def gather_packed(packed, group=None):
    """Exchange the flat per-field detection tensors of every rank with TWO collectives (sizes, then one padded byte
    payload) and return a list of PackedDetections, one per rank, in rank order. No per-image Python work."""
    from .detector import PackedDetections
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return [packed]
    world = dist.get_world_size(group)
    dev = packed.scores.device
    nimg = packed.num_images
    meta = torch.tensor(packed.triplet_off + packed.box_off, dtype=torch.int64, device=dev)       # 2*(B+1)
    # int64 fields first so every typed view of the byte payload stays 8-byte aligned
    parts = [packed.labels.contiguous().view(torch.uint8), packed.objects.contiguous().view(torch.uint8),
             packed.pairing.contiguous().view(torch.uint8), meta.view(torch.uint8),
             packed.scores.contiguous().view(torch.uint8), packed.boxes.contiguous().view(torch.uint8).reshape(-1)]
    payload = torch.cat(parts)
    sizes = torch.tensor([payload.numel(), nimg, packed.scores.numel(), packed.boxes.shape[0]], dtype=torch.int64, device=dev)
    all_sizes = torch.empty(world * 4, dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(all_sizes, sizes, group=group)
    all_sizes = all_sizes.view(world, 4).cpu().tolist()
    cap = (max(r[0] for r in all_sizes) + 15) // 16 * 16
    mine = torch.zeros(cap, dtype=torch.uint8, device=dev)
    mine[: payload.numel()] = payload
    gathered = torch.empty(world * cap, dtype=torch.uint8, device=dev)
    dist.all_gather_into_tensor(gathered, mine, group=group)
    out = []
    for r in range(world):
        nbytes, nb, m, nbox = all_sizes[r]
        buf = gathered[r * cap: r * cap + nbytes]
        o = 0

        def take(n_elems, dtype, esize):
            nonlocal o
            v = buf[o: o + n_elems * esize].view(dtype)
            o += n_elems * esize
            return v
        lb = take(m, torch.int64, 8)
        ob = take(m, torch.int64, 8)
        pr = take(2 * m, torch.int64, 8)
        mt_dev = take(2 * (nb + 1), torch.int64, 8)
        sc = take(m, torch.float32, 4)
        bx = take(nbox * 4, torch.float32, 4).view(nbox, 4)
        mt = mt_dev.cpu().tolist() if r != dist.get_rank(group) else packed.triplet_off + packed.box_off
        out.append(PackedDetections(sc, lb, ob, pr, bx, mt[: nb + 1], mt[nb + 1:], packed.size))
    return out


def gather_detections(dets: Sequence[Optional[dict]], group=None) -> List[dict]:
    """All ranks end up with the detections of every image, in global (rank-major, shard) order."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        c, f, i = pack_detections(dets)
        return unpack_detections(c, f, i)
    world = dist.get_world_size(group)
    counts, floats, ints = pack_detections(dets)
    dev = counts.device
    # 1) sizes: images, floats, int rows per rank
    sizes = torch.tensor([counts.shape[0], floats.numel(), ints.shape[0]], dtype=torch.int64, device=dev)
    all_sizes = [torch.zeros_like(sizes) for _ in range(world)]
    dist.all_gather(all_sizes, sizes, group=group)
    all_sizes = torch.stack(all_sizes).cpu()
    mx = all_sizes.max(0).values.tolist()

    def padded(t, n, shape_tail):
        buf = torch.zeros((n, *shape_tail), dtype=t.dtype, device=dev)
        buf[: t.shape[0]] = t
        return buf

    # 2) payload: one padded buffer per dtype
    bufs = []
    for t, n, tail in ((counts, mx[0], (2,)), (floats, mx[1], ()), (ints, mx[2], (4,))):
        mine = padded(t, n, tail)
        got = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(got, mine, group=group)
        bufs.append(got)
    out: List[dict] = []
    for r in range(world):
        nb, nf, ni = all_sizes[r].tolist()
        out.extend(unpack_detections(bufs[0][r][:nb], bufs[1][r][:nf], bufs[2][r][:ni]))
    return out
