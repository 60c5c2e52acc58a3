#!/usr/bin/env python
"""HOI-forward throughput benchmark (BASELINE.json metric: images/sec of the per-image HOI scoring forward).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]

A "step" is one pass of the hot path (region proposals -> detections: prior tokens, ViT-B/16+InsAdapter encoder,
RoIAlign/pair assembly, cache+text logits, prior scores + triplet emission) over ONE batch of synthetic images:
configs[1] of BASELINE.json — HICO-DET 117 verbs, batch 64, 8 human + 8 object boxes (120 pairs/image under the
reference's pairing rule), 4096x512 caches, bf16 tensor-core path.

  value     images/s with inputs resident in HBM (CUDA events; max over ranks; whole job)
  e2e       same metric through UPT.forward_from_proposals with HOST (pinned) inputs: H2D of the step's images /
            boxes and D2H of every detection tensor inside the timed region (wall clock between synchronisations)
  roofline  the dominant kernel (tcgen05 GEMM): algorithmic FLOPs of its launches / CUDA-event time, vs the measured
            bf16 peak in MEASURED_PEAKS.json
  cpu_baseline  the oracle port (reference algorithm, torch-CPU fp32, all host threads) on a bounded sample

`--impl reference` times the reference's CPU implementation of the path (oracle port; the reference is Python and
cannot travel to the GPU box) with all host threads on the same workload.
"""
from __future__ import annotations

import argparse
import gc
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

METRIC = "hoi_forward_images_per_sec"
UNIT = "images/s"
ENC_GFLOP_PER_IMG = 35.875
BOXES_H, BOXES_O = 8, 8


def workload_config(args, world):
    return {
        "workload": "HICO-DET HOI scoring forward (BASELINE configs[1]): ViT-B/16+InsAdapter 224^2, 117 verbs, "
                    f"{BOXES_H}h+{BOXES_O}o boxes = 120 pairs/img, {args.cache_rows}x512 caches (H,O,U,global,DINO) + text",
        "batch_per_gpu": args.batch, "global_batch": args.batch * world, "pairs_per_image": BOXES_H * (BOXES_H + BOXES_O - 1),
        "cache_rows": args.cache_rows, "num_classes": 117, "parallelism": f"image-sharded dp{world}",
        "l2": f"inputs rotate over {args.rotate} distinct batches ({args.rotate * args.batch * 3 * 224 * 224 * 4 / 1e6:.0f} MB "
              "> 126 MB L2); weights + activations per step >> L2",
        "dino_features": "supplied as input (SURVEY.md 8 row a8: stock ResNet-50 is outside the path)",
        "batches_in_flight": (f"{getattr(args, 'streams', 1)} (consecutive steps alternate over {getattr(args, 'streams', 1)} CUDA streams and "
                              "overlap on the GPU; kernel_breakdown / roofline are single-stream per-launch times)"),
        "collective": ("none (single GPU)" if world == 1 else
                       ("one NCCL all-gather of the sweep's accumulated detections, inside each timed region"
                        if getattr(args, "gather_every", 0) == 0 else "one NCCL all-gather of the step's detections per step")),
    }


class ClockSampler:
    """SM clock + clock-event (throttle) reasons sampled DURING the timed regions (B200_PROFILING.md clocks line).
    In-process NVML (nvidia_ml_py) on a thread: a looping `nvidia-smi -lms` child was measured to stall this process's
    kernel launches for ~4 ms per query, i.e. it perturbed the number it was there to qualify.  Falls back to one
    nvidia-smi query per second if NVML cannot be loaded."""

    _REASONS = (("hw_slowdown", "nvmlClocksEventReasonHwSlowdown", 0x8), ("hw_thermal_slowdown", "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
                ("sw_thermal_slowdown", "nvmlClocksEventReasonSwThermalSlowdown", 0x20), ("sw_power_cap", "nvmlClocksEventReasonSwPowerCap", 0x4))

    def __init__(self, index: int, period_s: float = 0.01):
        self.index, self.period = index, period_s
        self.sm, self.mx, self.reasons, self.stop_flag, self.thread, self.source = [], [], set(), False, None, None
        self.active = False     # samples are kept only while a timed region is running

    def _visible_index(self):
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        if vis:
            ids = [v.strip() for v in vis.split(",") if v.strip()]
            if self.index < len(ids) and ids[self.index].isdigit():
                return int(ids[self.index])
        return self.index

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            self.h = pynvml.nvmlDeviceGetHandleByIndex(self._visible_index())
            self.nv = pynvml
            self.source = "nvml"
            target = self._loop_nvml
        except Exception:
            self.source = "nvidia-smi"
            target = self._loop_smi
        self.thread = threading.Thread(target=target, daemon=True)
        self.thread.start()

    def _loop_nvml(self):
        nv = self.nv
        get_reasons = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or nv.nvmlDeviceGetCurrentClocksThrottleReasons
        while not self.stop_flag:
            if not self.active:
                time.sleep(self.period)
                continue
            try:
                self.sm.append(float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                self.mx.append(float(nv.nvmlDeviceGetMaxClockInfo(self.h, nv.NVML_CLOCK_SM)))
                bits = int(get_reasons(self.h))
                for name, _attr, bit in self._REASONS:
                    if bits & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(self.period)

    def _loop_smi(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        while not self.stop_flag:
            if not self.active:
                time.sleep(0.05)
                continue
            try:
                out = subprocess.run(["nvidia-smi", f"--id={self._visible_index()}", f"--query-gpu={q}", "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=10).stdout.strip().splitlines()
                r = [c.strip() for c in out[0].split(",")]
                self.sm.append(float(r[0])); self.mx.append(float(r[1]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[2:6]):
                    if v.lower().startswith("active"):
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(1.0)

    def stop(self):
        self.stop_flag = True
        if self.thread is not None:
            self.thread.join(timeout=15)
        sm = sorted(self.sm)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no clock samples"], "samples": 0, "source": self.source}
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": max(self.mx), "reasons": sorted(self.reasons), "samples": len(sm),
                "source": self.source}


def peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return d.get("bf16_tflops_sustained", d.get("bf16_tflops")), d.get("hbm_gbs"), "measured (MEASURED_PEAKS.json, sustained)"
    return 1400.0, 6650.0, "fallback (B200_PROFILING.md)"


# ------------------------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the oracle port on the host cores
# ------------------------------------------------------------------------------------------------------------------
def cpu_port_images_per_sec(batch: int, passes: int, cache_rows: int, warmup: int = 1):
    from hoigen_b200 import synthetic as S
    from oracle import hoi_forward_ref as O
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    enc = S.make_encoder_state(0)
    head = S.make_head_state(117, cache_rows)
    imgs = S.make_images(batch, seed=1)
    props = S.make_region_props(batch, BOXES_H, BOXES_O)
    dino = S.make_dino_features(batch)
    times = []
    with torch.no_grad():
        for i in range(warmup + passes):
            t0 = time.perf_counter()
            O.hoi_forward(imgs, props, dino, enc, head, roi_impl="torchvision")
            if i >= warmup:
                times.append(time.perf_counter() - t0)
    dt = sum(times) / len(times)
    return batch / dt, dt * 1e3, cores


def run_reference(args, rank, world):
    if rank != 0:
        return 0
    batch = 8
    value, ms, cores = cpu_port_images_per_sec(batch, max(args.steps, 1), args.cache_rows, warmup=min(args.warmup, 2))
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic (seeded random-init weights, images, boxes)",
        "config": workload_config(args, 1),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"each step = one batch of {batch} images of the same workload through the oracle port "
                                   "(reference algorithm, torch-CPU fp32 + torchvision roi_align, all host threads)"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


# ------------------------------------------------------------------------------------------------------------------
# B200 arm
# ------------------------------------------------------------------------------------------------------------------
def run_b200(args, rank, world, local_rank):
    import torch.distributed as dist
    from hoigen_b200 import _cabi, synthetic as S
    from hoigen_b200.detector import UPT
    from hoigen_b200.gather import gather_packed, gather_packed_begin, gather_packed_end, merge_packed

    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl b200 needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        # stdout carries exactly one JSON line: NCCL's version banner (this image exports NCCL_DEBUG=VERSION) is dropped,
        # any more verbose NCCL debugging the caller asked for goes to stderr
        if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=dev)
    _cabi.init(dev)
    B = args.batch
    model = UPT.from_state(S.make_encoder_state(0), S.make_head_state(117, args.cache_rows)).to(dev)
    model.pack_weights()
    model.clip_head.image_encoder.pack_weights()

    # ---- rotating input sets: host (pinned) + device resident copies ------------------------------------------------
    R = args.rotate
    host_imgs, host_props, host_dino = [], [], []
    for r in range(R):
        host_imgs.append(S.make_images(B, seed=1000 * rank + r + 1).pin_memory())
        props = [S.make_boxes(1000 * rank + 64 * r + b, BOXES_H, BOXES_O) for b in range(B)]
        host_props.append([{k: v.pin_memory() for k, v in p.items()} for p in props])
        host_dino.append(S.make_dino_features(B, seed=7 + r).pin_memory())
    dev_imgs = [t.to(dev) for t in host_imgs]
    dev_props = [[dict({k: v.to(dev) for k, v in p.items()}, n_human=BOXES_H) for p in ps] for ps in host_props]
    dev_dino = [t.to(dev) for t in host_dino]

    # consecutive steps alternate over `--streams` CUDA streams: the latency-bound head / prologue kernels of one batch
    # then overlap the GEMMs of the next instead of leaving most SMs idle
    streams = [torch.cuda.Stream(device=dev) for _ in range(args.streams)] if args.streams > 1 else [torch.cuda.current_stream(dev)]

    def launch_resident(i):
        r = i % R
        with torch.cuda.stream(streams[i % len(streams)]):
            return model.launch_from_proposals(dev_imgs[r], dev_props[r], dev_dino[r])

    def finish_resident(pend):
        dets = model.finish(pend)          # the path's device->host read (triplet offsets) + detection views
        if world > 1:
            exchange(dets.packed)
        return dets

    # The path's one collective: every rank ends up with every rank's detections.  One fixed-capacity all-gather per
    # step, enqueued without a host sync and read back one step later (so it never stalls the step launched ahead).
    gather_cap = B * 120 * 30 * 36 + B * (BOXES_H + BOXES_O) * 16
    pending_gather = []

    sweep = []

    def exchange(packed):
        if args.gather_every == 0:         # default: accumulate on the device, ONE exchange per sweep (= timed region)
            sweep.append(packed)
            return
        pending_gather.append(gather_packed_begin(packed, gather_cap, B))
        if len(pending_gather) > 1:
            gather_packed_end(pending_gather.pop(0))

    def drain_exchanges():
        while pending_gather:
            gather_packed_end(pending_gather.pop(0))
        if sweep:
            gather_packed(merge_packed(sweep))
            sweep.clear()

    def run_resident(first, count):
        """`count` complete steps; step i+1 is enqueued before step i is waited for, so the host's per-step work (layout,
        launches, result views) overlaps the GPU instead of leaving it idle.  Every step is launched AND finished here."""
        trace = os.environ.get("HOIGEN_BENCH_TRACE")
        pend = launch_resident(first)
        dets = None
        for i in range(count):
            ta = time.perf_counter()
            nxt = launch_resident(first + i + 1) if i + 1 < count else None
            tb = time.perf_counter()
            dets = finish_resident(pend)
            pend = nxt
            if nxt is None:
                drain_exchanges()
            if trace:
                print(f"[trace] step {first + i}: launch {1e3 * (tb - ta):.2f} ms, finish {1e3 * (time.perf_counter() - tb):.2f} ms",
                      file=sys.stderr, flush=True)
        return dets

    # ---- end-to-end step: HOST (pinned) inputs -> device -> detections -> HOST, H2D of step i+1 overlapped with step i --
    copy_stream = torch.cuda.Stream(device=dev)
    n_per = BOXES_H + BOXES_O
    host_packed = []
    for r in range(R):   # what a host-side caller holds: per-image proposals packed into three pinned arrays
        host_packed.append((torch.cat([p["boxes"] for p in host_props[r]]).pin_memory(),
                            torch.cat([p["scores"] for p in host_props[r]]).pin_memory(),
                            torch.cat([p["labels"] for p in host_props[r]]).pin_memory()))
    NSLOT = 3   # input slots: one being computed on, one launched ahead, one being uploaded
    dev_in = [dict(imgs=torch.empty_like(dev_imgs[0]), boxes=torch.empty(B * n_per, 4, device=dev),
                   scores=torch.empty(B * n_per, device=dev), labels=torch.empty(B * n_per, dtype=torch.int64, device=dev),
                   dino=torch.empty_like(dev_dino[0]), ev=torch.cuda.Event(), free=None) for _ in range(NSLOT)]
    cap_out = B * 120 * 30
    host_out = [dict(scores=torch.empty(cap_out, dtype=torch.float32).pin_memory(),
                     labels=torch.empty(cap_out, dtype=torch.int64).pin_memory(),
                     objects=torch.empty(cap_out, dtype=torch.int64).pin_memory(),
                     pairing=torch.empty(2 * cap_out, dtype=torch.int64).pin_memory(), ev=torch.cuda.Event(), keep=None)
                for _ in range(2)]
    d2h_stream = torch.cuda.Stream(device=dev)

    def upload(i):
        r, slot = i % R, dev_in[i % NSLOT]
        with torch.cuda.stream(copy_stream):
            if slot["free"] is not None:
                copy_stream.wait_event(slot["free"])     # the forward that last read this slot has finished
            slot["imgs"].copy_(host_imgs[r], non_blocking=True)
            slot["boxes"].copy_(host_packed[r][0], non_blocking=True)
            slot["scores"].copy_(host_packed[r][1], non_blocking=True)
            slot["labels"].copy_(host_packed[r][2], non_blocking=True)
            slot["dino"].copy_(host_dino[r], non_blocking=True)
            slot["ev"].record(copy_stream)
        return slot

    def launch_host(i):
        slot = dev_in[i % NSLOT]
        with torch.cuda.stream(streams[i % len(streams)]):
            torch.cuda.current_stream().wait_event(slot["ev"])
            pend = model.launch_packed(slot["imgs"], slot["boxes"], slot["scores"], slot["labels"], [n_per] * B, [BOXES_H] * B,
                                       slot["dino"])
        slot["free"] = pend.done
        return pend

    def finish_host(i, pend):
        """Wait for step i, then start the device->host copy of its detections on the D2H stream (it overlaps the next
        step's compute; the host buffer is double-buffered and waited for one step later)."""
        dets = model.finish(pend)
        pk = dets.packed
        if world > 1:
            exchange(pk)
        m = pk.scores.numel()
        ho = host_out[i % 2]
        if ho["keep"] is not None:
            ho["ev"].synchronize()                       # the copy that last used this host buffer has landed
        with torch.cuda.stream(d2h_stream):
            d2h_stream.wait_event(pk.done)
            ho["scores"][:m].copy_(pk.scores, non_blocking=True)
            ho["labels"][:m].copy_(pk.labels, non_blocking=True)
            ho["objects"][:m].copy_(pk.objects, non_blocking=True)
            ho["pairing"][: 2 * m].copy_(pk.pairing, non_blocking=True)
            ho["ev"].record(d2h_stream)
        ho["keep"] = dets                                # keeps the device tensors alive until the copy is done
        return m

    def run_host(first, count, pend):
        """`count` end-to-end steps starting at step `first` (already uploaded and launched as `pend`).  Per step: one
        upload (two steps ahead), one launch (one step ahead), one finish + D2H.  Returns the next pending step."""
        m = 0
        trace = os.environ.get("HOIGEN_BENCH_TRACE")
        evs = []
        for i in range(first, first + count):
            ta = time.perf_counter()
            if trace:
                evs.append(torch.cuda.Event(enable_timing=True))
                evs[-1].record()
            nxt = launch_host(i + 1)
            tb = time.perf_counter()
            m = finish_host(i, pend)
            tc = time.perf_counter()
            upload(i + 3)                                # its slot was read by step i, which has finished
            pend = nxt
            if trace:
                ms_ = torch.cuda.memory_stats(dev)
                print(f"[trace-e2e] step {i}: launch {1e3 * (tb - ta):.2f} ms, finish+d2h {1e3 * (tc - tb):.2f} ms, "
                      f"upload {1e3 * (time.perf_counter() - tc):.2f} ms segs {ms_['segment.all.allocated']} "
                      f"gc {[g['collections'] for g in gc.get_stats()]}", file=sys.stderr, flush=True)
        if trace and len(evs) > 2:
            torch.cuda.synchronize()
            print("[trace-e2e] GPU period between step starts (ms): " +
                  " ".join(f"{evs[k].elapsed_time(evs[k + 1]):.2f}" for k in range(len(evs) - 1)), file=sys.stderr, flush=True)
        return pend, m

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- device-resident timing -----------------------------------------------------------------------------------------
    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()                           # NVML is initialised here, outside the timed regions
    wdets = run_resident(0, args.warmup)
    if world > 1 and args.gather_every == 0:
        # warm the sweep-sized exchange too (NCCL channel setup / allocator growth for a K-step payload are one-time costs)
        gather_packed(merge_packed([wdets.packed] * args.steps))
    barrier()
    _cabi.profile(False)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    clocks.active = True
    e0.record()
    for st in streams:
        st.wait_event(e0)
    dets = run_resident(args.warmup, args.steps)
    for st in streams:
        torch.cuda.current_stream().wait_stream(st)
    e1.record()
    barrier()
    clocks.active = False
    launches = _cabi.launch_count()
    ms_step = max_over_ranks(e0.elapsed_time(e1) / args.steps)
    value = world * B / (ms_step * 1e-3)
    triplets = sum(int(d["scores"].numel()) for d in dets[:B])

    # ---- end-to-end timing with host buffers -----------------------------------------------------------------------------
    for i in range(3):
        upload(i)
    w_e2e = max(8, args.warmup)
    pend, m_out = run_host(0, w_e2e, launch_host(0))
    barrier()
    # The timer starts at a step boundary of the RUNNING pipeline, six untimed steps after that synchronisation: traced
    # (HOIGEN_BENCH_TRACE prints the caching allocator's segment count per step), the allocator grows by two segments on
    # the third launch after any full device synchronisation, whatever the warm-up length (step 11 after an 8-step
    # warm-up, step 27 after 24), and that cudaMalloc costs 1.5 ms normally but 20-160 ms in about one run in six - all of
    # it billed to this lockstep loop (one step in flight).  Starting without a synchronisation means step w_e2e + 6,
    # launched before t0, may still be running when the window opens, so the window holds AT LEAST `steps` whole steps of
    # device work plus all their copies (the final barrier drains the step launched ahead): a pessimistic boundary.
    pend, m_out = run_host(w_e2e, 6, pend)
    w_e2e += 6
    sweep.clear()                                # the timed sweep exchanges its own `steps` steps of detections, no warm-up ones
    clocks.active = True
    t0 = time.perf_counter()
    # K uploads, K launches, K finishes + K D2H copies; the barrier below waits for the last launched step and the copies
    pend, m_out = run_host(w_e2e, args.steps, pend)
    for ho in host_out:
        ho["ev"].synchronize()
    drain_exchanges()
    barrier()
    e2e_ms = max_over_ranks((time.perf_counter() - t0) * 1e3 / args.steps)
    clocks.active = False
    clk = clocks.stop() if rank == 0 else None
    h2d = host_imgs[0].numel() * 4 + sum(t.numel() * t.element_size() for t in host_packed[0]) + host_dino[0].numel() * 4
    d2h = m_out * (4 + 8 + 8 + 16) + (B + 1) * 4

    if os.environ.get("HOIGEN_BENCH_TRACE"):     # diagnostics: the same per-kernel profile, but of end-to-end steps
        torch.cuda.synchronize()
        _cabi.profile(True)
        pend, _ = run_host(w_e2e + args.steps, 2, pend)
        torch.cuda.synchronize()
        recs_e = _cabi.profile_read()
        _cabi.profile(False)
        agg_e = {}
        for tag, ms, fl, by, _t0 in recs_e:
            a = agg_e.setdefault(tag, [0, 0.0])
            a[0] += 1; a[1] += ms
        print("[trace-e2e] profiled kernels: %d launches, sum %.3f ms; per tag ms: %s" % (
            len(recs_e), sum(a[1] for a in agg_e.values()),
            " ".join(f"{t}={a[1] / a[0] * 1e3:.1f}us" for t, a in sorted(agg_e.items(), key=lambda kv: -kv[1][1])[:10])),
            file=sys.stderr, flush=True)
        ts = sorted((t0, ms, tag) for tag, ms, fl, by, t0 in recs_e)
        gaps = sorted(((ts[k + 1][0] - (ts[k][0] + ts[k][1])), ts[k][2], ts[k + 1][2]) for k in range(len(ts) - 1))[-6:]
        print("[trace-e2e] largest gaps (ms, after, before): " + "; ".join(f"{g:.3f} {a}->{b}" for g, a, b in gaps), file=sys.stderr, flush=True)
    # ---- opt-in folded-cache variant, reported NEXT TO the headline (which always runs the unfolded cache GEMMs) -------------
    folded = None
    if world == 1:
        torch.cuda.synchronize()
        model.fold_cache = True
        model.invalidate_packed()
        model.pack_weights()
        run_resident(0, 3)
        torch.cuda.synchronize()
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        f0.record()
        for st in streams:
            st.wait_event(f0)
        run_resident(3, args.steps)
        for st in streams:
            torch.cuda.current_stream().wait_stream(st)
        f1.record()
        torch.cuda.synchronize()
        fms = f0.elapsed_time(f1) / args.steps
        folded = {"ms_per_step": fms, "value": B / (fms * 1e-3), "unit": UNIT,
                  "note": "UPT(fold_cache=True): every linear cache contracted with its label matrix at pack time "
                          "(hoigen_score_pairs_folded); same detections (tests), NOT used for value / e2e above"}
        model.fold_cache = False
        model.invalidate_packed()
        model.pack_weights()
        torch.cuda.synchronize()
    # ---- per-kernel event profile of one step (separate from the timed regions) ---------------------------------------------
    _cabi.profile(True)
    for i in range(2):
        model.forward_from_proposals(dev_imgs[i % R], dev_props[i % R], dev_dino[i % R])
    recs = _cabi.profile_read()
    _cabi.profile(False)
    agg = {}
    for tag, ms, fl, by, _t0 in recs:
        a = agg.setdefault(tag, [0, 0.0, 0.0, 0.0])
        a[0] += 1; a[1] += ms; a[2] += fl; a[3] += by
    nprof = 2
    # the dominant kernel = the CTA-pair GEMM (tags gemm2_*: every encoder GEMM + the cache-affinity GEMMs)
    gemm_ms = sum(a[1] for t, a in agg.items() if t.startswith("gemm2_"))
    gemm_fl = sum(a[2] for t, a in agg.items() if t.startswith("gemm2_"))
    gemm_n = sum(a[0] for t, a in agg.items() if t.startswith("gemm2_"))
    allg_ms = sum(a[1] for t, a in agg.items() if t.startswith("gemm"))
    allg_fl = sum(a[2] for t, a in agg.items() if t.startswith("gemm"))
    total_ms = sum(a[1] for a in agg.values())
    peak_tf, peak_hbm, peak_src = peaks()
    achieved = gemm_fl / (gemm_ms * 1e-3) / 1e12 if gemm_ms > 0 else 0.0
    traffic = None
    tr = ROOT / "profiles" / "ncu_gemm_traffic.json"
    if tr.exists():
        traffic = json.loads(tr.read_text()).get("dram_bytes_per_launch")
    roofline = {"bound": "tensor", "kernel": "hoigen::gemm2_bf16_kernel (CTA-pair tcgen05/TMA GEMM; all its launches of a step)",
                "achieved": achieved, "peak": peak_tf, "unit": "TFLOP/s", "frac": achieved / peak_tf if peak_tf else None,
                "peak_source": peak_src, "traffic": traffic, "launches_per_step": gemm_n / nprof,
                "avg_launch_us": gemm_ms / gemm_n * 1e3 if gemm_n else None,
                "algorithmic_gflop_per_launch": gemm_fl / gemm_n / 1e9 if gemm_n else None,
                "share_of_step": gemm_ms / total_ms if total_ms else None,
                "all_gemm_kernels_tflops": allg_fl / (allg_ms * 1e-3) / 1e12 if allg_ms > 0 else None,
                "all_gemm_kernels_share_of_step": allg_ms / total_ms if total_ms else None}
    breakdown = {t: {"launches_per_step": a[0] / nprof, "ms_per_step": a[1] / nprof,
                     "tflops": (a[2] / (a[1] * 1e-3) / 1e12) if a[1] > 0 and a[2] > 0 else None,
                     "gbs": (a[3] / (a[1] * 1e-3) / 1e9) if a[1] > 0 and a[3] > 0 else None} for t, a in sorted(agg.items())}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0
    cpu_base = None
    if world == 1 and not args.no_cpu_baseline:
        cpu_passes = 24                                  # ~12 s of CPU work at ~17 images/s
        v, ms, cores = cpu_port_images_per_sec(8, cpu_passes, args.cache_rows)
        cpu_base = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                    "sample": f"{cpu_passes} timed passes (1 warm-up) of one batch of 8 images of the same workload "
                              f"({cpu_passes * 8} images, ~{cpu_passes * 8 / max(v, 1e-9):.0f} s) through the oracle port "
                              "(torch-CPU fp32 + torchvision roi_align, all host threads)"}
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16",
        "data": "synthetic (seeded random-init ViT-B/16+adapter weights, randn images, NMS-safe grid boxes)",
        "config": workload_config(args, world),
        "e2e": {"value": world * B / (e2e_ms * 1e-3), "unit": UNIT, "ms_per_step": e2e_ms, "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": d2h},
        "gpu_launches": launches,
        "clocks": clk,
        "roofline": roofline,
        "cpu_baseline": cpu_base,
        "encoder_tflops_effective": ENC_GFLOP_PER_IMG * B / (ms_step * 1e-3) / 1e3,
        "triplets_per_step": triplets,
        "kernel_breakdown": breakdown,
        "folded_cache_variant": folded,
    }
    print(json.dumps(line))
    out_dir = ROOT / "gpurun_out"
    if out_dir.exists():
        (out_dir / f"bench_detail_n{world}.json").write_text(json.dumps(line, indent=1))
        with open(out_dir / f"bench_timeline_n{world}.txt", "w") as f:   # launch timeline of the profiled steps (gaps = host stalls)
            prev_end = 0.0
            for tag, ms, fl, by, t0 in recs:
                f.write(f"{t0:10.4f} {ms:9.4f} gap={t0 - prev_end:8.4f} {tag}\n")
                prev_end = t0 + ms
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--cache-rows", type=int, default=4096)
    ap.add_argument("--rotate", type=int, default=4)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--streams", type=int, default=2,
                    help="CUDA streams that consecutive steps alternate over: 2 (default) = two batches overlap on the GPU, "
                         "1 = strictly one batch at a time")
    ap.add_argument("--gather-every", type=int, default=0,
                    help="N>1 only. 0 (default): detections accumulate on the device and are exchanged with ONE all-gather per "
                         "sweep (= per timed region, inside it); 1: one non-blocking fixed-capacity all-gather per step")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    if args.impl == "reference":
        return run_reference(args, rank, world)
    return run_b200(args, rank, world, local_rank)


if __name__ == "__main__":
    sys.exit(main())
