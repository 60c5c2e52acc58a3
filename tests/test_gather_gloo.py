"""CPU: the N>1 path (image sharding + variable-length detection gather) with world_size=2 over gloo."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from hoigen_b200.gather import gather_detections, pack_detections, shard_range, unpack_detections


def _fake_dets(rank, n_img):
    g = torch.Generator().manual_seed(100 + rank)
    out = []
    for b in range(n_img):
        m = int(torch.randint(0, 50, (1,), generator=g)) if (rank + b) % 3 else 0     # some images emit nothing
        n = int(torch.randint(2, 9, (1,), generator=g))
        out.append(dict(boxes=torch.rand(n, 4, generator=g), pairing=torch.randint(0, n, (2, m), generator=g),
                        scores=torch.rand(m, generator=g), labels=torch.randint(0, 117, (m,), generator=g),
                        objects=torch.randint(0, 80, (m,), generator=g), size=torch.tensor([224, 224])))
    return out


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        lo, hi = shard_range(7, rank, world)
        mine = _fake_dets(rank, hi - lo)
        allv = gather_detections(mine)
        expect = _fake_dets(0, shard_range(7, 0, world)[1]) + _fake_dets(1, 7 - shard_range(7, 0, world)[1])
        ok = len(allv) == 7
        for a, e in zip(allv, expect):
            for k in ("boxes", "pairing", "scores", "labels", "objects"):
                ok = ok and torch.equal(a[k], e[k]) and a[k].dtype == e[k].dtype
        q.put((rank, ok))
    finally:
        dist.destroy_process_group()


def _fake_packed(rank, n_img):
    from hoigen_b200.detector import PackedDetections
    dets = _fake_dets(rank, n_img)
    toff, boff = [0], [0]
    for d in dets:
        toff.append(toff[-1] + d["scores"].numel())
        boff.append(boff[-1] + d["boxes"].shape[0])
    pairing = torch.cat([d["pairing"].reshape(-1) for d in dets]) if dets else torch.zeros(0, dtype=torch.int64)
    return dets, PackedDetections(torch.cat([d["scores"] for d in dets]), torch.cat([d["labels"] for d in dets]),
                                  torch.cat([d["objects"] for d in dets]), pairing, torch.cat([d["boxes"] for d in dets]),
                                  toff, boff, (224, 224))


def _worker_packed(rank, world, port, q):
    from hoigen_b200.gather import gather_packed
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        n_img = 3 + rank      # ragged shards, odd triplet counts (alignment of the byte payload)
        _, mine = _fake_packed(rank, n_img)
        got = gather_packed(mine)
        ok = len(got) == world
        # the non-blocking form: two exchanges in flight, read back in order; fixed capacity agreed by all ranks
        from hoigen_b200.gather import gather_packed_begin, gather_packed_end
        h1 = gather_packed_begin(mine, 1 << 16, 8)
        h2 = gather_packed_begin(mine, 1 << 16, 8)
        for got2 in (gather_packed_end(h1), gather_packed_end(h2)):
            ok = ok and len(got2) == world
            for r in range(world):
                ok = ok and got2[r].triplet_off == got[r].triplet_off and got2[r].box_off == got[r].box_off
                for f in ("scores", "labels", "objects", "pairing", "boxes"):
                    ok = ok and torch.equal(getattr(got2[r], f), getattr(got[r], f))
        try:
            gather_packed_begin(mine, 64, 8)        # capacity too small: loud, before any collective is issued
            ok = False
        except ValueError:
            pass
        for r in range(world):
            ref_dets, _ = _fake_packed(r, 3 + r)
            ok = ok and got[r].num_images == 3 + r
            for b, e in enumerate(ref_dets):
                a = got[r].image(b)
                for k in ("boxes", "pairing", "scores", "labels", "objects"):
                    ok = ok and torch.equal(a[k], e[k]) and a[k].dtype == e[k].dtype
        q.put((rank, ok))
    finally:
        dist.destroy_process_group()


def test_gather_packed_world2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker_packed, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, True), (1, True)]


def test_merge_packed_preserves_every_image():
    """A sweep's batches merged into one record (what is exchanged once per sweep): per-image views are unchanged."""
    from hoigen_b200.gather import merge_packed
    dets_a, pa = _fake_packed(0, 3)
    dets_b, pb = _fake_packed(1, 4)
    dets_c, pc = _fake_packed(2, 1)
    merged = merge_packed([pa, None, pb, pc])
    assert merged.num_images == 8
    for b, e in enumerate(dets_a + dets_b + dets_c):
        a = merged.image(b)
        for k in ("boxes", "pairing", "scores", "labels", "objects"):
            assert torch.equal(a[k], e[k]) and a[k].dtype == e[k].dtype, (b, k)
    with pytest.raises(ValueError):
        merge_packed([None])


def test_shard_range_partitions():
    for total in (0, 1, 7, 512, 4097):
        for world in (1, 2, 8):
            spans = [shard_range(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert max(h - l for l, h in spans) - min(h - l for l, h in spans) <= 1


def test_pack_unpack_roundtrip_with_empty_and_none():
    dets = _fake_dets(0, 5) + [None]
    c, f, i = pack_detections(dets)
    back = unpack_detections(c, f, i)
    assert len(back) == 6 and back[5]["scores"].numel() == 0 and back[5]["pairing"].shape == (2, 0)
    for a, e in zip(back[:5], dets[:5]):
        for k in ("boxes", "pairing", "scores", "labels", "objects"):
            assert torch.equal(a[k], e[k])


def test_gather_world2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, True), (1, True)]


def test_wire_record_roundtrip_and_layout():
    """The compact wire record (9 B per triplet): torch pack -> torch unpack restores every field and dtype; the byte
    layout agrees with the C library's size function; a record that does not fit says so."""
    from hoigen_b200 import _cabi
    from hoigen_b200.gather import pack_wire_torch, unpack_wire_torch, wire_layout, wire_record_bytes
    dets, pk = _fake_packed(3, 6)
    m, nbox = pk.scores.numel(), pk.boxes.shape[0]
    cap = wire_record_bytes(8, m + 11, nbox + 3)
    assert cap == _cabi.load().hoigen_wire_record_bytes(8, m + 11, nbox + 3)
    lay = wire_layout(8, m, nbox)
    assert lay["end"] - lay["scores"] == 9 * m and lay["boxes"] % 16 == 0 and lay["labels"] % 2 == 0
    rec = torch.zeros(cap, dtype=torch.uint8)
    pack_wire_torch(pk, 8, cap, rec)
    back = unpack_wire_torch(rec, 8, (224, 224))
    assert back.triplet_off == pk.triplet_off and back.box_off == pk.box_off
    for b, e in enumerate(dets):
        a = back.image(b)
        for k in ("boxes", "pairing", "scores", "labels", "objects"):
            assert torch.equal(a[k], e[k]) and a[k].dtype == e[k].dtype, (b, k)
    small = torch.zeros(lay["end"] - 16, dtype=torch.uint8)
    pack_wire_torch(pk, 8, small.numel(), small)
    with pytest.raises(ValueError):
        unpack_wire_torch(small, 8, (224, 224))


def _worker_sweep(rank, world, port, q):
    from hoigen_b200.gather import SweepExchange
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        steps = 5
        ex = SweepExchange(world, 8, 8 * 50, 8 * 9, "cpu", max_steps=6)
        ok = True
        for sweep in range(2):                      # the exchanger is reusable across sweeps
            for s in range(steps):
                _, pk = _fake_packed(10 * rank + s + 100 * sweep, 3 + (rank + s) % 4)
                ex.add(pk)
            got = ex.finish()
            ok = ok and len(got) == world
            for r in range(world):
                ok = ok and len(got[r]) == steps
                for s in range(steps):
                    ref_dets, _ = _fake_packed(10 * r + s + 100 * sweep, 3 + (r + s) % 4)
                    ok = ok and got[r][s].num_images == len(ref_dets)
                    for b, e in enumerate(ref_dets):
                        a = got[r][s].image(b)
                        for k in ("boxes", "pairing", "scores", "labels", "objects"):
                            ok = ok and torch.equal(a[k], e[k]) and a[k].dtype == e[k].dtype
        q.put((rank, ok))
    finally:
        dist.destroy_process_group()


def test_sweep_exchange_world2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 33500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker_sweep, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, True), (1, True)]
