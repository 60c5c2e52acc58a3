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


# ----------------------------------------------------------------------------------------------------------------------
# Sweep exchange in the compact wire format (include/hoigen_b200.h "Multi-GPU"): 9 bytes per triplet instead of 36, one
# fixed-capacity all-gather per chunk of steps on a side stream, no host synchronisation until the sweep ends.
# ----------------------------------------------------------------------------------------------------------------------
WIRE_MAGIC = 0x57494F48


def _align(v: int, a: int) -> int:
    return (v + a - 1) // a * a


def wire_layout(max_images: int, m: int, nbox: int):
    """Byte offsets of the planes of one record (mirrors csrc/wire.cu::wire_layout)."""
    hdr_words = 4 + 2 * (max_images + 1)
    boxes = _align(hdr_words * 4, 16)
    scores = boxes + nbox * 16
    labels = scores + m * 4
    objects = labels + m * 2
    ph = objects + m
    po = ph + m
    return dict(hdr_words=hdr_words, boxes=boxes, scores=scores, labels=labels, objects=objects, ph=ph, po=po, end=po + m)


def wire_record_bytes(max_images: int, max_triplets: int, max_boxes: int) -> int:
    return _align(wire_layout(max_images, max_triplets, max_boxes)["end"], 16)


def pack_wire_torch(packed, max_images: int, cap_bytes: int, record: torch.Tensor) -> None:
    """Host / torch form of hoigen_pack_wire (CPU tensors: the gloo tests; also the format's executable description)."""
    nimg, m, nbox = packed.num_images, packed.scores.numel(), packed.boxes.shape[0]
    lay = wire_layout(max_images, m, nbox)
    hdr = torch.zeros(lay["hdr_words"], dtype=torch.int32)
    fits = lay["end"] <= cap_bytes
    hdr[:4] = torch.tensor([WIRE_MAGIC, nimg, m if fits else -1, nbox], dtype=torch.int64).to(torch.int32)
    hdr[4: 4 + nimg + 1] = torch.tensor(packed.triplet_off, dtype=torch.int32)
    hdr[4 + max_images + 1: 4 + max_images + 1 + nimg + 1] = torch.tensor(packed.box_off, dtype=torch.int32)
    record[: lay["hdr_words"] * 4] = hdr.view(torch.uint8).to(record.device)
    if not fits:
        return
    ph, po = [], []
    for b in range(nimg):
        s, e = packed.triplet_off[b], packed.triplet_off[b + 1]
        blk = packed.pairing[2 * s: 2 * e].view(2, e - s)
        ph.append(blk[0]); po.append(blk[1])
    cat = lambda xs: torch.cat(xs) if xs else torch.zeros(0, dtype=torch.int64, device=record.device)
    planes = [("boxes", packed.boxes.float().reshape(-1)), ("scores", packed.scores.float()),
              ("labels", packed.labels.to(torch.int16)), ("objects", packed.objects.to(torch.uint8)),
              ("ph", cat(ph).to(torch.uint8)), ("po", cat(po).to(torch.uint8))]
    for name, t in planes:
        raw = t.contiguous().view(torch.uint8).reshape(-1)
        record[lay[name]: lay[name] + raw.numel()] = raw


def unpack_wire_torch(record: torch.Tensor, max_images: int, size):
    """Host / torch form of hoigen_unpack_wire for ONE record -> PackedDetections (int64 indices, [2][M_b] pairing blocks)."""
    from .detector import PackedDetections
    hdr_words = 4 + 2 * (max_images + 1)
    hdr = record[: hdr_words * 4].contiguous().view(torch.int32).tolist()
    if (hdr[0] & 0xFFFFFFFF) != WIRE_MAGIC:
        raise ValueError("not a detection wire record")
    nimg, m, nbox = hdr[1], hdr[2], hdr[3]
    if m < 0:
        raise ValueError("the sender's detections did not fit the record capacity")
    toff, boff = hdr[4: 4 + nimg + 1], hdr[4 + max_images + 1: 4 + max_images + 1 + nimg + 1]
    lay = wire_layout(max_images, m, nbox)
    take = lambda name, n, dt, es: record[lay[name]: lay[name] + n * es].contiguous().view(dt)
    boxes = take("boxes", nbox * 4, torch.float32, 4).view(nbox, 4)
    scores = take("scores", m, torch.float32, 4)
    labels = take("labels", m, torch.int16, 2).to(torch.int64) & 0xFFFF
    objects = take("objects", m, torch.uint8, 1).to(torch.int64)
    ph, po = take("ph", m, torch.uint8, 1).to(torch.int64), take("po", m, torch.uint8, 1).to(torch.int64)
    pairing = torch.empty(2 * m, dtype=torch.int64, device=record.device)
    for b in range(nimg):
        s, e = toff[b], toff[b + 1]
        pairing[2 * s: s + e] = ph[s:e]
        pairing[s + e: 2 * e] = po[s:e]
    return PackedDetections(scores, labels, objects, pairing, boxes, toff, boff, size)


class _SweepRecord:
    """One (rank, step) record of a finished sweep: PackedDetections whose per-field tensors are views into the sweep-wide
    output tensors, cut on first access (a sweep at 8 ranks x 32 steps holds 256 records x 5 fields)."""

    def __init__(self, bufs, tb, m, bb, nbox, triplet_off, box_off, size, done):
        self._bufs, self._tb, self._m, self._bb, self._nbox = bufs, tb, m, bb, nbox
        self.triplet_off, self.box_off, self.size, self.done = triplet_off, box_off, tuple(size), done

    scores = property(lambda self: self._bufs[0][self._tb: self._tb + self._m])
    labels = property(lambda self: self._bufs[1][self._tb: self._tb + self._m])
    objects = property(lambda self: self._bufs[2][self._tb: self._tb + self._m])
    pairing = property(lambda self: self._bufs[3][2 * self._tb: 2 * (self._tb + self._m)])
    boxes = property(lambda self: self._bufs[4][self._bb: self._bb + self._nbox])

    @property
    def num_images(self) -> int:
        return len(self.triplet_off) - 1

    def image(self, b: int) -> dict:
        from .detector import PackedDetections
        return PackedDetections.image(self, b)


class SweepExchange:
    """All ranks end up with every rank's detections of a sweep (an evaluation pass over a dataset shard).

    `add(packed)` after every finished step packs the step's detections into the compact wire record ON THE DEVICE, on a
    side stream (hoigen_pack_wire reads the forward's own device-side offsets).  Transport of the records:

      "p2p"         (CUDA, world > 1, default when torch's symmetric memory can be set up): every rank owns a receive
                    buffer [rank][slot] mapped into all peers (NVLink / NVSwitch peer memory); a record is PUSHED straight
                    into every peer's buffer with copy-engine peer copies — no SM is used, no rank waits for another, so
                    the exchange overlaps the next steps' compute without disturbing it (an NCCL all-gather kernel per
                    chunk was measured to DOUBLE the step time at N = 2: its CTAs hold SMs while they spin for the peer,
                    which breaks the co-residency the persistent GEMMs are scheduled for).  `finish()` = one barrier.
      "collective"  ONE all-gather of the sweep's records at `finish()` (NCCL; gloo in the CPU tests), after the compute.

    No host synchronisation happens until `finish()`, which reads all headers back once, widens every record to the
    reference's dtypes (hoigen_unpack_wire) and returns [rank][step] PackedDetections.  A sweep holds at most `max_steps`
    steps (`full` tells when to `finish()`); every rank must add the same number of steps per sweep."""

    def __init__(self, world: int, max_images: int, max_triplets: int, max_boxes: int, device, max_steps: int = 32,
                 group=None, transport: str = "auto"):
        self.world, self.max_images, self.max_steps, self.group = world, max_images, max(1, max_steps), group
        self.dev = torch.device(device)
        self.cap = wire_record_bytes(max_images, max_triplets, max_boxes)
        self.hdr_bytes = (4 + 2 * (max_images + 1)) * 4
        self.cuda = self.dev.type == "cuda"
        self.side = torch.cuda.Stream(device=self.dev) if self.cuda else None
        self.rank = dist.get_rank(group) if (world > 1 and dist.is_initialized()) else 0
        self.transport, self.symm = "collective", None
        nbytes = world * self.max_steps * self.cap
        if self.cuda and world > 1 and transport in ("auto", "p2p"):
            try:
                import torch.distributed._symmetric_memory as symm_mem
                grp = group if group is not None else dist.group.WORLD
                self.recv = symm_mem.empty(nbytes, dtype=torch.uint8, device=self.dev)
                self.symm = symm_mem.rendezvous(self.recv, grp.group_name if hasattr(grp, "group_name") else grp)
                self.peers = [self.symm.get_buffer(r, (nbytes,), torch.uint8) for r in range(world)]
                self.transport = "p2p"
            except Exception as e:          # no peer mapping on this box / build: the collective form still works
                if transport == "p2p":
                    raise
                self.symm, self.why_not_p2p = None, f"{type(e).__name__}: {e}"
        if self.transport == "collective":
            self.recv = torch.empty(nbytes, dtype=torch.uint8, device=self.dev)
        # this rank's records of the sweep (p2p: the own slice of the receive buffer)
        self.mine = (self.recv[self.rank * self.max_steps * self.cap: (self.rank + 1) * self.max_steps * self.cap]
                     if self.transport == "p2p" else torch.empty(self.max_steps * self.cap, dtype=torch.uint8, device=self.dev))
        # pinned landing / staging buffers for finish(), allocated ONCE (a pinned allocation per call is a cudaHostAlloc: it
        # synchronises the device and was measured at 10-30 ms for a first-seen size)
        if self.cuda:
            self.hdr_host = torch.empty(world, self.max_steps, self.hdr_bytes, dtype=torch.uint8).pin_memory()
            self.bases_host = torch.empty(world, self.max_steps, 2, dtype=torch.int64).pin_memory()
            self.bases_dev = torch.empty(world, self.max_steps, 2, dtype=torch.int64, device=self.dev)
        self._reset()

    def _reset(self):
        self.n = 0
        self.keep = []          # the steps' own tensors: alive until their records have been built
        self.size = None

    @property
    def full(self) -> bool:
        return self.n >= self.max_steps

    def add(self, packed, pend=None) -> None:
        import os, time
        t_add = time.perf_counter() if os.environ.get("HOIGEN_GATHER_TRACE") else None
        try:
            return self._add(packed, pend)
        finally:
            if t_add is not None:
                self._t_add = getattr(self, "_t_add", 0.0) + time.perf_counter() - t_add

    def _add(self, packed, pend=None) -> None:
        if self.full:
            raise RuntimeError(f"SweepExchange holds {self.max_steps} steps per sweep: call finish() first")
        self.size = packed.size
        rec = self.mine[self.n * self.cap: (self.n + 1) * self.cap]
        if not self.cuda:
            rec[: self.hdr_bytes].zero_()
            pack_wire_torch(packed, self.max_images, self.cap, rec)
            self.n += 1
            return
        from . import _cabi
        done = getattr(packed, "done", None)
        with torch.cuda.stream(self.side):
            if done is not None:
                self.side.wait_event(done)
            else:
                self.side.wait_stream(torch.cuda.current_stream(self.dev))
            img_off = packed.img_off_dev if pend is None else pend.img_off       # device-side int32 offsets of the forward
            box_off = packed.box_off_dev if pend is None else pend.d_box_off
            _cabi.call("hoigen_pack_wire", packed.scores.data_ptr(), packed.labels.data_ptr(), packed.objects.data_ptr(),
                       packed.pairing.data_ptr(), packed.boxes.data_ptr(), img_off.data_ptr(), box_off.data_ptr(),
                       packed.num_images, self.max_images, self.cap, rec.data_ptr())
            if self.transport == "p2p":
                # the record's size is known on the host (the forward's offsets came back with finish()): push exactly
                # that many bytes into slot [rank][n] of every peer's receive buffer — copy engines over NVLink, no SMs
                nb = min(self.cap, _align(wire_layout(self.max_images, packed.scores.numel(), packed.boxes.shape[0])["end"], 16))
                off = (self.rank * self.max_steps + self.n) * self.cap
                for r in range(self.world):
                    if r != self.rank:
                        self.peers[r][off: off + nb].copy_(rec[:nb], non_blocking=True)
        self.keep.append((packed, pend))
        self.n += 1

    def finish(self):
        """-> [rank][step] PackedDetections of everything added since the last finish()."""
        import os, sys, time
        if not os.environ.get("HOIGEN_GATHER_TRACE"):
            return self._finish()
        if self.cuda:
            torch.cuda.synchronize(self.dev)
        t0 = time.perf_counter()
        n = self.n
        out = self._finish()
        if self.cuda:
            torch.cuda.synchronize(self.dev)
        if self.rank == 0:
            print(f"[exchange] finish of {n} steps: {1e3 * (time.perf_counter() - t0):.2f} ms (device idle before it); host time in add() "
                  f"since the last report {1e3 * getattr(self, '_t_add', 0.0):.2f} ms; transport {self.transport}", file=sys.stderr, flush=True)
        self._t_add = 0.0
        return out

    def _finish(self):
        from .detector import PackedDetections
        W, S, cap, hb, n = self.world, self.max_steps, self.cap, self.hdr_bytes, self.n
        out = [[] for _ in range(W)]
        if n == 0:
            return out
        if not self.cuda:
            if W > 1:
                gathered = torch.empty(W * n * cap, dtype=torch.uint8)
                dist.all_gather_into_tensor(gathered, self.mine[: n * cap].contiguous(), group=self.group)
                g = gathered.view(W, n, cap)
            else:
                g = self.mine[: n * cap].view(1, n, cap)
            for r in range(W):
                for s in range(n):
                    out[r].append(unpack_wire_torch(g[r, s], self.max_images, self.size))
            self._reset()
            return out
        from . import _cabi
        with torch.cuda.stream(self.side):
            if self.transport == "p2p":
                self.symm.barrier()                    # every rank's pushes (stream-ordered before its barrier) have landed
                g = self.recv.view(W, S, cap)
            elif W > 1:
                gathered = torch.empty(W * n * cap, dtype=torch.uint8, device=self.dev)
                dist.all_gather_into_tensor(gathered, self.mine[: n * cap], group=self.group)
                g = gathered.view(W, n, cap)
            else:
                g = self.mine.view(1, S, cap)
            # ONE read-back: every record's header
            self.hdr_host[:, :n].copy_(g[:, :n, :hb], non_blocking=True)
            ev = torch.cuda.Event()
            ev.record()
        ev.synchronize()
        H = self.hdr_host[:, :n].numpy().view("int32").reshape(W, n, -1)
        mi = self.max_images
        if (H[:, :, 0] != WIRE_MAGIC).any():
            r, s = [int(v[0]) for v in (H[:, :, 0] != WIRE_MAGIC).nonzero()]
            raise ValueError(f"rank {r} slot {s}: no record arrived (ranks must add the same number of steps per sweep)")
        if (H[:, :, 2] < 0).any():
            raise ValueError(f"rank {int((H[:, :, 2] < 0).nonzero()[0][0])}: a step's detections did not fit the record capacity ({cap} bytes)")
        # host side of the unpack, vectorised: at 8 ranks x 20 steps a per-record Python loop over tensor slices cost ~3 ms
        import numpy as np
        m_arr, nb_arr = H[:, :, 2].astype(np.int64), H[:, :, 3].astype(np.int64)
        tb_arr = (np.cumsum(m_arr.ravel()) - m_arr.ravel()).reshape(W, n)      # triplet / box base of record (r, s) in the
        bb_arr = (np.cumsum(nb_arr.ravel()) - nb_arr.ravel()).reshape(W, n)    # sweep-wide output tensors, (rank, step) order
        m_tot, b_tot = int(m_arr.sum()), int(nb_arr.sum())
        rows = g.shape[1]                                         # slots per rank in the buffer being unpacked
        with torch.cuda.stream(self.side):
            scores = torch.empty(max(m_tot, 1), dtype=torch.float32, device=self.dev)
            labels = torch.empty(max(m_tot, 1), dtype=torch.int64, device=self.dev)
            objects = torch.empty(max(m_tot, 1), dtype=torch.int64, device=self.dev)
            pairing = torch.empty(max(2 * m_tot, 2), dtype=torch.int64, device=self.dev)
            boxes = torch.empty(max(b_tot, 1), 4, dtype=torch.float32, device=self.dev)
            bases_np = self.bases_host.numpy()
            bases_np[...] = -1
            bases_np[:, :n, 0] = tb_arr
            bases_np[:, :n, 1] = bb_arr
            self.bases_dev.copy_(self.bases_host, non_blocking=True)
            bases_d = self.bases_dev
            base_ptr = g.data_ptr()
            for r in range(W):                                    # one launch per rank: its slots are contiguous
                _cabi.call("hoigen_unpack_wire", base_ptr + r * rows * cap, min(rows, n), cap, mi, bases_d[r].data_ptr(), scores.data_ptr(),
                           labels.data_ptr(), objects.data_ptr(), pairing.data_ptr(), boxes.data_ptr())
            if self.transport == "p2p":
                # headers cleared + a second barrier: no peer may push the next sweep's records while this one is being read
                self.recv.view(W, S, cap)[:, :, :hb].zero_()
                self.symm.barrier()
            done = torch.cuda.Event()
            done.record()
        torch.cuda.current_stream(self.dev).wait_event(done)
        Hl = H.tolist()
        bufs = (scores, labels, objects, pairing, boxes)
        for r in range(W):
            row_out = out[r]
            for s in range(n):
                h = Hl[r][s]
                nimg = h[1]
                row_out.append(_SweepRecord(bufs, int(tb_arr[r, s]), h[2], int(bb_arr[r, s]), h[3], h[4: 4 + nimg + 1],
                                            h[4 + mi + 1: 4 + mi + 1 + nimg + 1], self.size, done))
        self._reset()
        return out
