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


_PINNED_HEADERS = {}      # (world, hdr_words) -> [ring of pinned int64 buffers, next index]


def _pinned_header(world: int, hdr_words: int) -> torch.Tensor:
    """Pinned landing buffer for the gathered headers, from a small ring: a fresh pinned allocation per exchange would
    call cudaHostAlloc, which synchronises the device (and with it every forward the caller has in flight)."""
    ring = _PINNED_HEADERS.setdefault((world, hdr_words), [[], 0])
    if len(ring[0]) < 4:
        ring[0].append(torch.empty(world, hdr_words, dtype=torch.int64).pin_memory())
        return ring[0][-1]
    ring[1] = (ring[1] + 1) % 4
    return ring[0][ring[1]]


class _PendingGather:
    """An exchange that has been enqueued (gather_packed_begin) but not read back (gather_packed_end)."""


def _payload_parts(packed, dev):
    meta = torch.tensor(packed.triplet_off + packed.box_off, dtype=torch.int64).to(dev, non_blocking=True)   # 2*(B+1)
    # int64 fields first so every typed view of the byte payload stays 8-byte aligned
    return [packed.labels.contiguous().view(torch.uint8), packed.objects.contiguous().view(torch.uint8),
            packed.pairing.contiguous().view(torch.uint8), meta.view(torch.uint8),
            packed.scores.contiguous().view(torch.uint8), packed.boxes.contiguous().view(torch.uint8).reshape(-1)]


def _unpack(buf, nb, m, nbox, meta_host, size):
    from .detector import PackedDetections
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
    mt = meta_host if meta_host is not None else mt_dev.cpu().tolist()
    return PackedDetections(sc, lb, ob, pr, bx, mt[: nb + 1], mt[nb + 1:], size)


def gather_packed_begin(packed, cap_bytes: int, max_images: int, group=None) -> _PendingGather:
    """Enqueue the exchange of `packed` as ONE all-gather of a fixed-capacity byte record per rank, on the current stream,
    without any host synchronisation: [header: nbytes, nimg, m, nbox, triplet_off.., box_off..][fields].  `cap_bytes` /
    `max_images` are bounds every rank agrees on (e.g. from the batch geometry).  The header block comes back through a
    pinned buffer + event, so `gather_packed_end` waits for exactly this exchange and nothing enqueued after it."""
    world = dist.get_world_size(group)
    dev = packed.scores.device
    nimg = packed.num_images
    if nimg > max_images:
        raise ValueError(f"{nimg} images > max_images = {max_images}")
    hdr_words = 4 + 2 * (max_images + 1)
    parts = _payload_parts(packed, dev)
    nbytes = sum(p.numel() for p in parts)
    cap = (hdr_words * 8 + cap_bytes + 15) // 16 * 16
    if hdr_words * 8 + nbytes > cap:
        raise ValueError(f"detections need {nbytes} bytes, cap_bytes = {cap_bytes}")
    hdr = torch.zeros(hdr_words, dtype=torch.int64)
    hdr[:4] = torch.tensor([nbytes, nimg, packed.scores.numel(), packed.boxes.shape[0]])
    hdr[4: 4 + 2 * (nimg + 1)] = torch.tensor(packed.triplet_off + packed.box_off)
    mine = torch.empty(cap, dtype=torch.uint8, device=dev)
    mine[: hdr_words * 8] = hdr.to(dev, non_blocking=True).view(torch.uint8)
    torch.cat(parts, out=mine[hdr_words * 8: hdr_words * 8 + nbytes])
    gathered = torch.empty(world * cap, dtype=torch.uint8, device=dev)
    dist.all_gather_into_tensor(gathered, mine, group=group)
    h = _PendingGather()
    h.gathered, h.cap, h.hdr_words, h.world, h.size, h.group, h.local = gathered, cap, hdr_words, world, packed.size, group, packed
    if dev.type == "cuda":
        h.hdr_host = _pinned_header(world, hdr_words)          # at most 3 exchanges may be in flight
        h.hdr_host.copy_(gathered.view(world, cap)[:, : hdr_words * 8].contiguous().view(torch.int64).view(world, hdr_words),
                         non_blocking=True)
        h.event = torch.cuda.Event()
        h.event.record()
    else:
        h.hdr_host = gathered.view(world, cap)[:, : hdr_words * 8].contiguous().view(torch.int64).view(world, hdr_words).clone()
        h.event = None
    return h


def gather_packed_end(h: _PendingGather):
    """Wait for an exchange started by gather_packed_begin; returns one PackedDetections per rank, in rank order."""
    if h.event is not None:
        h.event.synchronize()
    out = []
    for r in range(h.world):
        row = h.hdr_host[r].tolist()
        nbytes, nb, m, nbox = row[:4]
        meta = row[4: 4 + 2 * (nb + 1)]
        start = r * h.cap + h.hdr_words * 8
        out.append(_unpack(h.gathered[start: start + nbytes], nb, m, nbox, meta, h.size))
    return out


def merge_packed(batches):
    """Concatenate the PackedDetections of consecutive batches of one rank into one record (image order preserved).
    An evaluation sweep accumulates its batches on the device and exchanges them ONCE (`gather_packed(merge_packed(..))`):
    the reference computes its metrics after the whole sweep, and a per-batch collective would make every step run at
    the pace of the slowest rank."""
    from .detector import PackedDetections
    batches = [b for b in batches if b is not None]
    if not batches:
        raise ValueError("merge_packed: nothing to merge")
    toff, boff = [0], [0]
    for b in batches:
        toff += [toff[-1] + v for v in b.triplet_off[1:]]
        boff += [boff[-1] + v for v in b.box_off[1:]]
    cat = lambda name: torch.cat([getattr(b, name) for b in batches])
    merged = PackedDetections(cat("scores"), cat("labels"), cat("objects"), cat("pairing"), cat("boxes"), toff, boff,
                              batches[0].size)
    merged.done = getattr(batches[-1], "done", None)
    return merged


def gather_packed(packed, group=None):
    """Exchange the flat per-field detection tensors of every rank with TWO collectives (sizes, then one padded byte
    payload) and return a list of PackedDetections, one per rank, in rank order. No per-image Python work.  Blocking
    convenience form; a serving loop that keeps forwards in flight uses gather_packed_begin / gather_packed_end."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return [packed]
    import os, sys, time
    trace = os.environ.get("HOIGEN_GATHER_TRACE")
    marks = []

    def mark(name):
        if trace:
            if packed.scores.is_cuda:
                torch.cuda.synchronize()
            marks.append((name, time.perf_counter()))
    mark("start")
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    dev = packed.scores.device
    nimg = packed.num_images
    parts = _payload_parts(packed, dev)
    nbytes = sum(p.numel() for p in parts)
    sizes = torch.tensor([nbytes, nimg, packed.scores.numel(), packed.boxes.shape[0]], dtype=torch.int64).to(dev)
    all_sizes = torch.empty(world * 4, dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(all_sizes, sizes, group=group)
    all_sizes = all_sizes.view(world, 4).cpu().tolist()
    mark("sizes")
    cap = (max(r[0] for r in all_sizes) + 15) // 16 * 16
    gathered = torch.empty(world * cap, dtype=torch.uint8, device=dev)
    mine = gathered[rank * cap: (rank + 1) * cap]          # in place: this rank's record is written straight into its slot
    o = 0
    for p in parts:
        mine[o: o + p.numel()].copy_(p)
        o += p.numel()
    mark("pack")
    dist.all_gather_into_tensor(gathered, mine, group=group)
    mark("all_gather")
    # one read-back for every remote rank's offsets instead of one per rank
    metas, spans = [], []
    for r in range(world):
        nb_r, nimg_r, m_r, _ = all_sizes[r]
        if r != rank:
            start = r * cap + m_r * 8 * 4
            metas.append(gathered[start: start + 2 * (nimg_r + 1) * 8])
    meta_host = torch.cat(metas).view(torch.int64).cpu().tolist() if metas else []
    out, mo = [], 0
    for r in range(world):
        nbytes_r, nb, m, nbox = all_sizes[r]
        if r == rank:
            own = packed.triplet_off + packed.box_off
        else:
            own = meta_host[mo: mo + 2 * (nb + 1)]
            mo += 2 * (nb + 1)
        out.append(_unpack(gathered[r * cap: r * cap + nbytes_r], nb, m, nbox, own, packed.size))
    mark("unpack")
    if trace and rank == 0:
        print("[gather] " + " ".join(f"{n}={1e3 * (t - marks[i][1]):.2f}ms" for i, (n, t) in enumerate(marks[1:])) +
              f" payload={nbytes / 1e6:.1f}MB", file=sys.stderr, flush=True)
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
