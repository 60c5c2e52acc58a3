#!/usr/bin/env python
"""HOI-forward throughput benchmark (BASELINE.json metric: images/sec of the per-image HOI scoring forward).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--config 2|3|4|5] [--impl b200|reference|reference-gpu]

A "step" is one pass of the hot path (region proposals -> detections: prior tokens, ViT-B/16+InsAdapter encoder,
RoIAlign/pair assembly, cache+text logits, prior scores + triplet emission) over ONE batch of synthetic images.  The
default workload is BASELINE.json configs[1] (SURVEY.md 8d "Config 2"): HICO-DET 117 verbs, batch 64, 8 human + 8
object boxes (120 pairs/image under the reference's pairing rule), 4096x512 caches, bf16 tensor-core path.  `--config`
selects the other BASELINE configurations (SURVEY 8d numbering): 3 = uc0 zero-shot 16384-row cache with generator-made
rows, 4 = V-COCO 24 actions, batch 128, 16h+16o boxes (496 pairs), 5 = 600-triplet classifier, batch 512 per GPU.

  value     images/s with inputs resident in HBM (CUDA events; max over ranks; whole job)
  e2e       same metric through UPT.launch_packed / UPT.finish with HOST (pinned) inputs: H2D of the step's images /
            boxes and D2H of every detection tensor inside the timed region (wall clock between synchronisations)
  sustained the resident loop again for >= 3 s, with its own clock record (the 20-step value is a burst number)
  roofline  the dominant kernel (tcgen05 GEMM): algorithmic FLOPs of its launches / CUDA-event time, vs the measured
            bf16 peak in MEASURED_PEAKS.json
  cpu_baseline        the reference's CPU implementation on a bounded sample: the UNMODIFIED reference staged under
                      oracle/_ref (`kind: "reference"`), else the oracle port (`kind: "port"`)
  gpu_torch_baseline  the UNMODIFIED reference on the same B200, same inputs: fp32 as shipped and under
                      torch.autocast(bf16) (BASELINE configs[1] "vs reference torch path on the same GPU inputs")
  full_with_dino_r50  the step with the ResNet-50 DINO branch (U:1616-1618) run per step by the module: stock fp32 torchvision,
                      the cuDNN bf16 graph, and this repo's own convolution kernels (UPT.accelerate_dino)

`--impl reference` times the reference's CPU implementation of the path with all host threads on the same workload;
`--impl reference-gpu` is the same-GPU torch comparator as a stand-alone run (what gpu_torch_baseline spawns).
"""
from __future__ import annotations

import argparse
import gc
import json
import math
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

# SURVEY.md 8d numbering ("Config 2" = BASELINE.json configs[1], the one the metric is quoted on)
WORKLOADS = {
    2: dict(name="BASELINE configs[1]: HICO-DET 117 verbs, bf16, batch 64", C=117, N=4096, B=64, nh=8, no=8, head="plain",
            dataset="hicodet", max_instances=15, ref_args={}),
    3: dict(name="BASELINE configs[2]: UC zero-shot (uc0), 16384x512 caches ending in generator-synthesised unseen-class rows",
            C=117, N=16384, B=64, nh=8, no=8, head="uc0", dataset="hicodet", max_instances=15,
            ref_args=dict(zs=True, zs_type="uc0")),
    4: dict(name="BASELINE configs[3]: V-COCO 24 actions, batch 128, 16h+16o boxes", C=24, N=4096, B=128, nh=16, no=16,
            head="plain", dataset="vcoco", max_instances=16, ref_args=dict(max_instances=16, cache=True, eval=False)),
    5: dict(name="BASELINE configs[4]: 600-triplet HICO classifier, batch 512 per GPU, detection gather", C=600, N=4096, B=512,
            nh=8, no=8, head="plain", dataset="hicodet", max_instances=15, ref_args=dict(generate_feature=False)),
}


def workload(args):
    w = dict(WORKLOADS[args.config])
    if args.batch:
        w["B"] = args.batch
    if args.cache_rows:
        w["N"] = args.cache_rows
    w["n"] = w["nh"] + w["no"]
    w["K"] = w["nh"] * (w["n"] - 1)
    return w


def flops_per_image(w):
    """SURVEY.md 8d algorithmic FLOPs: encoder (35.875 GFLOP with 16 prior tokens; the adapter's cross-attention and K/V
    projection scale with the token count) and the scoring chain."""
    n, K, N, C = w["n"], w["K"], w["N"], w["C"]
    enc = 35.875e9 + 12 * (2 * 2 * 197 * 64 * (n - 16) + 2 * 64 * 128 * (n - 16))
    score = 3 * 2 * K * 512 * N + 3 * 2 * K * N * C + 2 * K * 512 * C + (2 * 512 * N + 2 * N * C) + (2 * 2048 * N + 2 * N * C)
    return enc, score


def workload_config(args, world, w):
    return {
        "workload": f"{w['name']}: ViT-B/16+InsAdapter 224^2, {w['C']} classes, {w['nh']}h+{w['no']}o boxes = {w['K']} pairs/img, "
                    f"{w['N']}x512 caches (H,O,U,global,DINO) + text",
        "survey_config": args.config, "batch_per_gpu": w["B"], "global_batch": w["B"] * world, "pairs_per_image": w["K"],
        "cache_rows": w["N"], "num_classes": w["C"], "parallelism": f"image-sharded dp{world}",
        "l2": f"inputs rotate over {args.rotate} distinct batches ({args.rotate * w['B'] * 3 * 224 * 224 * 4 / 1e6:.0f} MB "
              "> 126 MB L2); weights + activations per step >> L2",
        "dino_features": "supplied as input (SURVEY.md 8 row a8: stock ResNet-50 is outside the path); full_with_dino_r50 "
                         "reports the step with the stock R50 run per step",
        "batches_in_flight": (f"{getattr(args, 'streams', 1)} (consecutive steps alternate over {getattr(args, 'streams', 1)} CUDA streams and "
                              "overlap on the GPU; kernel_breakdown / roofline are single-stream per-launch times)"),
        "launch_ahead": (f"{getattr(args, 'ahead', 1)} steps are enqueued ahead of the one being finished (every step is launched and "
                         "finished inside its timed region)"),
        "collective": ("none (single GPU)" if world == 1 else
                       "all ranks receive all detections: per step a compact wire record (9 B/triplet) pushed into every peer's buffer over "
                       "NVLink peer memory (copy engines), one barrier + unpack per sweep, inside each timed region"),
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
        self.power = []
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
                self.power.append(float(nv.nvmlDeviceGetPowerUsage(self.h)) / 1e3)
                bits = int(get_reasons(self.h))
                for name, _attr, bit in self._REASONS:
                    if bits & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(self.period)

    def sample_now(self):
        """One synchronous NVML sample on the calling thread (the queued GPU work is still running): the 20-step window is
        only ~75 ms long and the sampling thread can be starved of the GIL by the launch loop."""
        if self.source != "nvml" or not self.active:
            return
        nv = self.nv
        get_reasons = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or nv.nvmlDeviceGetCurrentClocksThrottleReasons
        try:
            self.sm.append(float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
            self.mx.append(float(nv.nvmlDeviceGetMaxClockInfo(self.h, nv.NVML_CLOCK_SM)))
            self.power.append(float(nv.nvmlDeviceGetPowerUsage(self.h)) / 1e3)
            bits = int(get_reasons(self.h))
            for name, _attr, bit in self._REASONS:
                if bits & bit:
                    self.reasons.add(name)
        except Exception:
            pass

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

    def snapshot(self, reset: bool = True):
        """Summary of the samples taken since the last snapshot."""
        sm = sorted(self.sm)
        if not sm:
            out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no clock samples"], "samples": 0, "source": self.source}
        else:
            out = {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": max(self.mx), "reasons": sorted(self.reasons), "samples": len(sm),
                   "source": self.source}
            if self.power:
                out["power_w_max"] = max(self.power)
        if reset:
            self.sm, self.mx, self.power, self.reasons = [], [], [], set()
        return out

    def stop(self):
        self.stop_flag = True
        if self.thread is not None:
            self.thread.join(timeout=15)
        return self.snapshot(reset=False)


def peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return (d.get("bf16_tflops_sustained", d.get("bf16_tflops")), d.get("bf16_tflops"), d.get("hbm_gbs"),
                "measured (MEASURED_PEAKS.json)")
    return 1400.0, 1590.0, 6650.0, "fallback (B200_PROFILING.md)"


def make_state(w):
    from hoigen_b200 import synthetic as S
    enc = S.make_encoder_state(0)
    head = (S.make_head_state_uc0(w["N"]) if w["head"] == "uc0"
            else S.make_head_state(w["C"], w["N"], seed=2, max_instances=w["max_instances"]))
    return enc, head


# ------------------------------------------------------------------------------------------------------------------
# reference legs: the UNMODIFIED reference (oracle/_ref, staged by __graft_entry__.build()) or the oracle port
# ------------------------------------------------------------------------------------------------------------------
def _build_reference(w, force_cpu):
    from oracle import ref_harness as RH
    if not RH.available():
        return None
    upt, pp = RH.build_reference_upt(w["C"], w["dataset"], force_cpu=force_cpu, **w["ref_args"])
    enc, head = make_state(w)
    RH.load_synthetic_state(upt, enc, head)
    return RH, upt, pp


def reference_cpu_images_per_sec(w, batch: int, passes: int, warmup: int = 1):
    """-> (images/s, ms per pass, threads, kind).  kind 'reference' = the unmodified reference modules run through their
    own UPT.forward (stub DETR feeding the synthetic boxes through the real prepare_region_proposals); 'port' = the
    oracle restatement, when oracle/_ref is not staged."""
    from hoigen_b200 import synthetic as S
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    imgs = S.make_images(batch, seed=1)
    props = S.make_region_props(batch, w["nh"], w["no"])
    dino = S.make_dino_features(batch)
    built = _build_reference(w, force_cpu=True)
    times = []
    if built is not None:
        RH, upt, pp = built
        kind = "reference"
        run = lambda: RH.run_reference(upt, pp, imgs, props, dino)
    else:
        from oracle import hoi_forward_ref as O
        enc, head = make_state(w)
        kind = "port"
        run = lambda: O.hoi_forward(imgs, props, dino, enc, head, roi_impl="torchvision")
    with torch.no_grad():
        for i in range(warmup + passes):
            t0 = time.perf_counter()
            run()
            if i >= warmup:
                times.append(time.perf_counter() - t0)
    dt = sum(times) / len(times)
    return batch / dt, dt * 1e3, cores, kind


def run_reference(args, rank, world):
    """--impl reference: the reference's own CPU implementation of the path, all host threads, same workload; each step a
    bounded sample (one batch of `ref_batch` images).  Rank 0 alone runs it."""
    if rank != 0:
        return 0
    w = workload(args)
    batch = args.ref_batch
    value, ms, cores, kind = reference_cpu_images_per_sec(w, batch, max(args.steps, 1), warmup=min(args.warmup, 2))
    what = ("the UNMODIFIED reference (oracle/_ref: build_detector + UPT.forward, stub DETR -> real prepare_region_proposals)"
            if kind == "reference" else "the oracle port (oracle/_ref not staged)")
    cfg = workload_config(args, 1, w)
    cfg["reference_batch_per_step"] = batch
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic (seeded random-init weights, images, boxes)",
        "config": cfg,
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind,
                         "sample": f"each step = one batch of {batch} images of the same workload through {what}, torch-CPU fp32, "
                                   f"{cores} host threads"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


def run_reference_gpu(args):
    """--impl reference-gpu: the UNMODIFIED reference on cuda:0 with the same inputs — (i) as shipped (fp32, no autocast:
    U:1612 is commented out), (ii) under torch.autocast(bf16).  CUDA events, whole UPT.forward per step."""
    w = workload(args)
    if not torch.cuda.is_available():
        print(json.dumps({"impl": "reference-gpu", "unavailable": "no CUDA device"}))
        return 0
    built = _build_reference(w, force_cpu=False)
    if built is None:
        print(json.dumps({"impl": "reference-gpu", "unavailable": "oracle/_ref not staged (run __graft_entry__.build() where /root/reference exists)"}))
        return 0
    from hoigen_b200 import synthetic as S
    RH, upt, pp = built
    dev = torch.device("cuda:0")
    upt = upt.to(dev)
    for name in ("sample_lens_H", "sample_lens_O", "sample_lens_U", "dino_sample_len", "global_sample_len", "object_embedding",
                 "dino_cache_values", "clip_cache_values", "origin_text_embeddings"):
        t = getattr(upt, name, None)
        if torch.is_tensor(t):
            setattr(upt, name, t.to(dev))
    B = min(w["B"], args.ref_gpu_batch)
    imgs = S.make_images(B, seed=1).to(dev)
    props = [{k: v.to(dev) for k, v in p.items()} for p in S.make_region_props(B, w["nh"], w["no"])]
    dino = S.make_dino_features(B).to(dev)
    out = {"impl": "reference-gpu", "batch": B, "unit": UNIT,
           "what": "UNMODIFIED reference UPT.forward (oracle/_ref) on the same B200, same seeded inputs; stub DETR, real "
                   "prepare_region_proposals / get_prior / image_encoder / compute_roi_embeddings / postprocessing"}
    for mode in ("fp32_as_shipped", "bf16_autocast"):
        ctx = torch.autocast("cuda", dtype=torch.bfloat16) if mode == "bf16_autocast" else torch.autocast("cuda", enabled=False)
        times = []
        try:
            with ctx:
                for i in range(args.warmup + args.steps):
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                    dets = RH.run_reference(upt, pp, imgs, props, dino)
                    e1.record()
                    torch.cuda.synchronize()
                    if i >= args.warmup:
                        times.append(e0.elapsed_time(e1))
            times.sort()
            med = times[len(times) // 2]
            out[mode] = {"value": B / (med * 1e-3), "ms_per_step": med, "steps": len(times),
                         "triplets": int(sum(d["scores"].numel() for d in dets))}
        except Exception as e:   # reported, not fatal: this leg is a comparator
            out[mode] = {"unavailable": f"{type(e).__name__}: {e}"[:300]}
    print(json.dumps(out))
    return 0


def spawn_leg(argv, timeout):
    """Run a reference leg of this script in its own process (the harness monkey-patches `.cuda()` / chdir()s) and parse the
    single JSON line it prints."""
    try:
        r = subprocess.run([sys.executable, str(ROOT / "bench.py"), *argv], capture_output=True, text=True, timeout=timeout,
                           env={k: v for k, v in os.environ.items() if k not in ("RANK", "WORLD_SIZE", "LOCAL_RANK")})
        for ln in reversed(r.stdout.strip().splitlines()):
            if ln.startswith("{"):
                return json.loads(ln)
        return {"unavailable": f"no JSON from {' '.join(argv)} (rc {r.returncode}): {r.stderr[-300:]}"}
    except Exception as e:
        return {"unavailable": f"{type(e).__name__}: {e}"[:300]}


# ------------------------------------------------------------------------------------------------------------------
# B200 arm
# ------------------------------------------------------------------------------------------------------------------
def run_b200(args, rank, world, local_rank):
    import torch.distributed as dist
    from hoigen_b200 import _cabi, synthetic as S
    from hoigen_b200.detector import UPT
    from hoigen_b200.gather import SweepExchange

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
    w = workload(args)
    B, NH, NO, n_per = w["B"], w["nh"], w["no"], w["n"]
    enc_state, head_state = make_state(w)
    model = UPT.from_state(enc_state, head_state).to(dev)
    p_packed, _ = model.pack_weights()
    model.clip_head.image_encoder.pack_weights()
    max_row_len = p_packed["max_row_len"]

    # ---- rotating input sets: host (pinned) + device resident copies ------------------------------------------------
    R = args.rotate
    host_imgs, host_packed, host_dino = [], [], []
    for r in range(R):
        host_imgs.append(S.make_images(B, seed=1000 * rank + r + 1).pin_memory())
        props = [S.make_boxes(1000 * rank + B * r + b, NH, NO) for b in range(B)]
        # what a host-side caller holds: per-image proposals packed into three pinned arrays
        host_packed.append((torch.cat([p["boxes"] for p in props]).pin_memory(), torch.cat([p["scores"] for p in props]).pin_memory(),
                            torch.cat([p["labels"] for p in props]).pin_memory()))
        host_dino.append(S.make_dino_features(B, seed=7 + r).pin_memory())
    dev_imgs = [t.to(dev) for t in host_imgs]
    dev_packed = [tuple(t.to(dev) for t in hp) for hp in host_packed]
    dev_dino = [t.to(dev) for t in host_dino]
    n_list, nh_list = [n_per] * B, [NH] * B

    # consecutive steps alternate over `--streams` CUDA streams: the latency-bound head / prologue kernels of one batch
    # then overlap the GEMMs of the next instead of leaving most SMs idle
    streams = [torch.cuda.Stream(device=dev) for _ in range(args.streams)] if args.streams > 1 else [torch.cuda.current_stream(dev)]

    # "full" figure: the DINO ResNet-50 branch (U:1616-1618) run by the module itself instead of supplied features
    dino_state = {"inline": False}

    def launch_resident(i):
        r = i % R
        st = streams[i % len(streams)]
        with torch.cuda.stream(st):
            bx, sc, lb = dev_packed[r]
            return model.launch_packed(dev_imgs[r], bx, sc, lb, n_list, nh_list, None if dino_state["inline"] else dev_dino[r])

    # The path's one exchange (N > 1): every rank ends up with every rank's detections.  Each step's detections are packed
    # on a side stream into the compact wire record (9 B per triplet instead of 36) and pushed into every peer's receive
    # buffer over NVLink peer memory (copy engines, no SM, no rank waits for another); no host synchronisation inside the
    # loop; drained (one barrier, headers read, records widened back to int64) inside the timed region.
    exchange = (SweepExchange(world, B, w["K"] * max_row_len * B, n_per * B, dev, max_steps=max(args.steps, args.gather_every))
                if world > 1 else None)

    def exchange_add(packed, pend):
        if exchange.full:                  # sweeps longer than the exchanger's capacity (the sustained block): drain first
            exchange.finish()
        exchange.add(packed, pend)

    def finish_resident(pend):
        dets = model.finish(pend)          # the path's device->host read (triplet offsets) + detection views
        if exchange is not None:
            exchange_add(dets.packed, pend)
        return dets

    exch_stats = {"finish_ms": []}

    def drain_exchanges():
        if exchange is not None:
            t_f = time.perf_counter()
            out = exchange.finish()
            exch_stats["finish_ms"].append((time.perf_counter() - t_f) * 1e3)
            return out
        return None

    def run_resident(first, count):
        """`count` complete steps; steps i+1 .. i+A (--ahead, default 1) are enqueued before step i is waited for, so the
        host's per-step work (layout, launches, result views) overlaps the GPU.  Every step is launched AND finished here."""
        trace = os.environ.get("HOIGEN_BENCH_TRACE")
        A = max(1, args.ahead)
        queue = [launch_resident(first + j) for j in range(min(A, count))]
        dets = None
        for i in range(count):
            ta = time.perf_counter()
            if i + A < count:
                queue.append(launch_resident(first + i + A))
            tb = time.perf_counter()
            dets = finish_resident(queue.pop(0))
            if i + 1 == count:
                drain_exchanges()
            if trace:
                print(f"[trace] step {first + i}: launch {1e3 * (tb - ta):.2f} ms, finish {1e3 * (time.perf_counter() - tb):.2f} ms",
                      file=sys.stderr, flush=True)
        return dets

    def timed_resident(first, count):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for st in streams:
            st.wait_event(e0)
        dets = run_resident(first, count)
        for st in streams:
            torch.cuda.current_stream().wait_stream(st)
        e1.record()
        return e0, e1, dets

    # ---- end-to-end step: HOST (pinned) inputs -> device -> detections -> HOST, H2D of step i+1 overlapped with step i --
    copy_stream = torch.cuda.Stream(device=dev)
    NSLOT = max(1, args.ahead) + 2   # input slots: one being computed on, --ahead launched ahead, one being uploaded
    dev_in = [dict(imgs=torch.empty_like(dev_imgs[0]), boxes=torch.empty(B * n_per, 4, device=dev),
                   scores=torch.empty(B * n_per, device=dev), labels=torch.empty(B * n_per, dtype=torch.int64, device=dev),
                   dino=torch.empty_like(dev_dino[0]), ev=torch.cuda.Event(), free=None) for _ in range(NSLOT)]
    cap_out = B * w["K"] * max_row_len
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
            pend = model.launch_packed(slot["imgs"], slot["boxes"], slot["scores"], slot["labels"], n_list, nh_list, slot["dino"])
        slot["free"] = pend.done
        return pend

    def finish_host(i, pend):
        """Wait for step i, then start the device->host copy of its detections on the D2H stream (it overlaps the next
        step's compute; the host buffer is double-buffered and waited for one step later)."""
        dets = model.finish(pend)
        pk = dets.packed
        if exchange is not None:
            exchange_add(pk, pend)
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

    def run_host(first, count, queue):
        """`count` end-to-end steps starting at step `first`; `queue` holds the already launched steps first .. first+A-1
        (A = --ahead).  Per step: one upload (A + 2 steps ahead), one launch (A steps ahead), one finish + D2H.  Returns
        the queue of still-pending steps."""
        m = 0
        A = max(1, args.ahead)
        trace = os.environ.get("HOIGEN_BENCH_TRACE")
        for i in range(first, first + count):
            ta = time.perf_counter()
            queue.append(launch_host(i + A))
            tb = time.perf_counter()
            m = finish_host(i, queue.pop(0))
            tc = time.perf_counter()
            upload(i + A + 2)                            # its slot was read by step i, which has finished
            if trace:
                ms_ = torch.cuda.memory_stats(dev)
                print(f"[trace-e2e] step {i}: launch {1e3 * (tb - ta):.2f} ms, finish+d2h {1e3 * (tc - tb):.2f} ms, "
                      f"upload {1e3 * (time.perf_counter() - tc):.2f} ms segs {ms_['segment.all.allocated']} "
                      f"gc {[g['collections'] for g in gc.get_stats()]}", file=sys.stderr, flush=True)
        return queue, m

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
    run_resident(0, args.warmup)
    if world > 1:
        run_resident(0, args.steps)              # a full-length sweep: the exchange's sweep-sized allocations are one-time costs
    barrier()
    _cabi.profile(False)
    clocks.active = True
    e0, e1, dets = timed_resident(args.warmup, args.steps)
    if rank == 0:
        clocks.sample_now()                      # the last steps are still running on the GPU
    barrier()
    clocks.active = False
    launches = _cabi.launch_count()
    own_ms = e0.elapsed_time(e1) / args.steps
    per_rank_ms = None
    if world > 1:      # every rank's own step time: separates GPU-to-GPU variation from the cost of the exchange
        t_all = torch.zeros(world, device=dev, dtype=torch.float64)
        t_all[rank] = own_ms
        dist.all_reduce(t_all)
        per_rank_ms = [round(v, 4) for v in t_all.tolist()]
    timed_finish_ms = exch_stats["finish_ms"][-1] if exch_stats["finish_ms"] else None
    ms_step = max_over_ranks(own_ms)
    value = world * B / (ms_step * 1e-3)
    triplets = sum(int(d["scores"].numel()) for d in dets[:B])
    clk = clocks.snapshot() if rank == 0 else None

    # ---- sustained: the same loop for >= args.sustain_s seconds (the 20-step window above is a burst measurement) ---------
    sustained = None
    if args.sustain_s > 0:
        n_sus = max(args.steps, int(math.ceil(args.sustain_s * 1e3 / ms_step)))
        clocks.active = True
        s0, s1, _ = timed_resident(0, n_sus)
        if rank == 0:
            clocks.sample_now()
        barrier()
        clocks.active = False
        sus_ms = max_over_ranks(s0.elapsed_time(s1) / n_sus)
        sustained = {"steps": n_sus, "seconds": sus_ms * n_sus * 1e-3, "ms_per_step": sus_ms, "value": world * B / (sus_ms * 1e-3),
                     "unit": UNIT, "clocks": clocks.snapshot() if rank == 0 else None}

    # ---- end-to-end timing with host buffers -----------------------------------------------------------------------------
    for i in range(max(1, args.ahead) + 2):
        upload(i)
    w_e2e = max(8, args.warmup)
    pend, m_out = run_host(0, w_e2e, [launch_host(j) for j in range(max(1, args.ahead))])
    barrier()
    # The timer starts at a step boundary of the RUNNING pipeline, six untimed steps after that synchronisation: traced
    # (HOIGEN_BENCH_TRACE prints the caching allocator's segment count per step), the allocator grows by two segments on
    # the third launch after any full device synchronisation, whatever the warm-up length, and that cudaMalloc costs 1.5 ms
    # normally but 20-160 ms in about one run in six - all of it billed to this lockstep loop (one step in flight).
    # Starting without a synchronisation means step w_e2e + 6, launched before t0, may still be running when the window
    # opens, so the window holds AT LEAST `steps` whole steps of device work plus all their copies (the final barrier
    # drains the step launched ahead): a pessimistic boundary.
    pend, m_out = run_host(w_e2e, 6, pend)
    w_e2e += 6
    if exchange is not None:
        exchange.finish()                        # the timed sweep exchanges its own `steps` steps of detections
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
    if rank == 0:
        clocks.snapshot()
    h2d = host_imgs[0].numel() * 4 + sum(t.numel() * t.element_size() for t in host_packed[0]) + host_dino[0].numel() * 4
    d2h = m_out * (4 + 8 + 8 + 16) + (B + 1) * 4
    for p_ in pend:
        model.finish(p_)
    torch.cuda.synchronize()

    def quick_resident(steps):
        run_resident(0, 3)
        torch.cuda.synchronize()
        f0, f1, _ = timed_resident(3, steps)
        torch.cuda.synchronize()
        return f0.elapsed_time(f1) / steps

    # ---- opt-in folded-cache variant, reported NEXT TO the headline (which always runs the unfolded cache GEMMs) -------------
    folded = None
    full_dino = None
    if world == 1 and not args.no_variants:
        model.fold_cache = True
        model.invalidate_packed()
        model.pack_weights()
        fms = quick_resident(args.steps)
        folded = {"ms_per_step": fms, "value": B / (fms * 1e-3), "unit": UNIT,
                  "note": "UPT(fold_cache=True): every linear cache contracted with its label matrix at pack time "
                          "(hoigen_score_pairs_folded); same detections (tests), NOT used for value / e2e above"}
        model.fold_cache = False
        model.invalidate_packed()
        model.pack_weights()
        torch.cuda.synchronize()
        # ---- "full" step: the DINO ResNet-50 branch (U:1616-1618) run per step by the module (features NOT supplied) --------
        try:
            import torchvision
            r50 = torchvision.models.resnet50(weights=None)
            r50.fc = torch.nn.Identity()
            model.dino_model = r50.to(dev).eval()
            dino_state["inline"] = True
            full_dino = {"what": "same resident loop, but the module runs its DINO branch itself (torchvision ResNet-50, fc = Identity, random "
                                 "init, + L2 normalise: U:1616-1618) for every step; cuDNN, outside the four kernel groups (SURVEY 8 row a8)"}
            fms = quick_resident(args.steps)
            full_dino["stock_fp32"] = {"ms_per_step": fms, "value": B / (fms * 1e-3), "unit": UNIT,
                                       "note": "the injected module exactly as the reference runs it (fp32, eager)"}
            model.accelerate_dino(engine="cudnn")
            fms = quick_resident(args.steps)
            full_dino["accelerate_dino_cudnn"] = {"ms_per_step": fms, "value": B / (fms * 1e-3), "unit": UNIT,
                                                  "note": "UPT.accelerate_dino(engine='cudnn'): BatchNorms folded, bf16 channels-last, one CUDA "
                                                          "graph per batch size and stream (library tuning; bf16-accurate features)"}
            model.accelerate_dino(engine="kernels")
            fms = quick_resident(args.steps)
            full_dino["accelerate_dino_kernels"] = {"ms_per_step": fms, "value": B / (fms * 1e-3), "unit": UNIT,
                                                    "note": "UPT.accelerate_dino(engine='kernels'): the branch on this repo's own kernels -- every "
                                                            "convolution on the tcgen05 GEMM (3x3 as an implicit GEMM over haloed NHWC rows), fused "
                                                            "tensor-core stem, one C call per batch (hoigen_b200/dino.py KernelDinoR50; opt-in, "
                                                            "bf16-accurate features)"}
        except Exception as e:
            full_dino = dict(full_dino or {}, unavailable=f"{type(e).__name__}: {e}"[:300])
        dino_state["inline"] = False
        object.__setattr__(model, "_fast_dino", None)
        model.dino_model = None
        torch.cuda.synchronize()

    # ---- per-kernel event profile of one step (separate from the timed regions) ---------------------------------------------
    _cabi.profile(True)
    nprof = 2
    for i in range(nprof):
        bx, sc, lb = dev_packed[i % R]
        model.finish(model.launch_packed(dev_imgs[i % R], bx, sc, lb, n_list, nh_list, dev_dino[i % R]))
    recs = _cabi.profile_read()
    _cabi.profile(False)
    agg = {}
    for tag, ms, fl, by, _t0 in recs:
        a = agg.setdefault(tag, [0, 0.0, 0.0, 0.0])
        a[0] += 1; a[1] += ms; a[2] += fl; a[3] += by
    # the dominant kernel = the CTA-pair GEMM (tags gemm2_*: every encoder GEMM + the cache-affinity GEMMs)
    gemm_ms = sum(a[1] for t, a in agg.items() if t.startswith("gemm2_"))
    gemm_fl = sum(a[2] for t, a in agg.items() if t.startswith("gemm2_"))
    gemm_n = sum(a[0] for t, a in agg.items() if t.startswith("gemm2_"))
    allg_ms = sum(a[1] for t, a in agg.items() if t.startswith("gemm"))
    allg_fl = sum(a[2] for t, a in agg.items() if t.startswith("gemm"))
    total_ms = sum(a[1] for a in agg.values())
    peak_sus, peak_burst, peak_hbm, peak_src = peaks()
    achieved = gemm_fl / (gemm_ms * 1e-3) / 1e12 if gemm_ms > 0 else 0.0
    traffic = None
    tr = ROOT / "profiles" / "ncu_gemm_traffic.json"
    if tr.exists():
        traffic = json.loads(tr.read_text()).get("dram_bytes_per_launch")
    # the per-launch events come from a short single-stream profile pass at burst clocks: the like-for-like denominator is
    # the burst peak; the fraction of the sustained peak is printed beside it
    roofline = {"bound": "tensor", "kernel": "hoigen::gemm2_bf16_kernel (CTA-pair tcgen05/TMA GEMM; all its launches of a step)",
                "achieved": achieved, "peak": peak_burst, "unit": "TFLOP/s", "frac": achieved / peak_burst if peak_burst else None,
                "peak_source": peak_src + ", bf16 burst (kernel timed alone, per launch)",
                "frac_of_sustained_peak": achieved / peak_sus if peak_sus else None, "peak_sustained": peak_sus,
                "traffic": traffic, "traffic_source": "profiles/ncu_gemm_traffic.json (ncu --set full capture, mean over the GEMM launches of a layer)",
                "launches_per_step": gemm_n / nprof,
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
    clocks.stop()
    cfg_flag = ["--config", str(args.config)] + (["--batch", str(args.batch)] if args.batch else []) + \
               (["--cache-rows", str(args.cache_rows)] if args.cache_rows else [])
    cpu_base = None
    if world == 1 and not args.no_cpu_baseline:
        # ~10-25 s of CPU work on the box's host cores, in its own process (the harness neutralises the reference's
        # hard-coded .cuda() calls there)
        ref = spawn_leg(["--impl", "reference", "--steps", "24" if args.config == 2 else "6", "--warmup", "1",
                         "--ref-batch", "8", *cfg_flag], timeout=900)
        cpu_base = ref.get("cpu_baseline") or {"unavailable": ref.get("unavailable", "no cpu_baseline in the reference leg")}
    gpu_base = None
    if world == 1 and not args.no_gpu_baseline:
        torch.cuda.empty_cache()
        gpu_base = spawn_leg(["--impl", "reference-gpu", "--steps", "5", "--warmup", "2", *cfg_flag], timeout=900)
    enc_fl, score_fl = flops_per_image(w)
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16",
        "data": "synthetic (seeded random-init ViT-B/16+adapter weights, randn images, NMS-safe grid boxes)",
        "config": workload_config(args, world, w),
        "e2e": {"value": world * B / (e2e_ms * 1e-3), "unit": UNIT, "ms_per_step": e2e_ms, "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": d2h},
        "gpu_launches": launches,
        "clocks": clk,
        "sustained": sustained,
        "roofline": roofline,
        "cpu_baseline": cpu_base,
        "gpu_torch_baseline": gpu_base,
        "encoder_tflops_effective": enc_fl * B / (ms_step * 1e-3) / 1e12,
        "encoder_frac_of_sustained_peak": enc_fl * B / (ms_step * 1e-3) / 1e12 / peak_sus if peak_sus else None,
        "algorithmic_gflop_per_image": {"encoder": enc_fl / 1e9, "scoring": score_fl / 1e9},
        "triplets_per_step": triplets,
        "kernel_breakdown": breakdown,
        "folded_cache_variant": folded,
        "full_with_dino_r50": full_dino,
    }
    if exchange is not None:
        line["detection_exchange"] = {"transport": exchange.transport, "record_capacity_bytes": exchange.cap,
                                      "per_rank_ms_per_step": per_rank_ms,
                                      "rank0_host_ms_in_finish_of_timed_sweep": timed_finish_ms,
                                      "bytes_per_triplet": 9, "sweep_capacity_steps": exchange.max_steps,
                                      "why_not_p2p": getattr(exchange, "why_not_p2p", None)}
    print(json.dumps(line))
    out_dir = ROOT / "gpurun_out"
    if out_dir.exists():
        (out_dir / f"bench_detail_c{args.config}_n{world}.json").write_text(json.dumps(line, indent=1))
        with open(out_dir / f"bench_timeline_c{args.config}_n{world}.txt", "w") as f:   # launch timeline of the profiled steps (gaps = host stalls)
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
    ap.add_argument("--impl", default="b200", choices=["b200", "reference", "reference-gpu"])
    ap.add_argument("--config", type=int, default=2, choices=sorted(WORKLOADS),
                    help="SURVEY.md 8d configuration: 2 (default) = BASELINE configs[1]; 3 = uc0 16k cache; 4 = V-COCO batch 128; "
                         "5 = 600 triplets, batch 512 per GPU")
    ap.add_argument("--batch", type=int, default=0, help="override the configuration's batch per GPU")
    ap.add_argument("--cache-rows", type=int, default=0, help="override the configuration's cache rows")
    ap.add_argument("--rotate", type=int, default=0, help="distinct input batches the steps rotate over (default: enough to exceed L2)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-gpu-baseline", action="store_true")
    ap.add_argument("--no-variants", action="store_true", help="skip the folded-cache and full-with-DINO side measurements")
    ap.add_argument("--sustain-s", type=float, default=3.0, help="length of the sustained block in seconds (0 = skip)")
    ap.add_argument("--ref-batch", type=int, default=8, help="--impl reference: images per step (bounded CPU sample)")
    ap.add_argument("--ref-gpu-batch", type=int, default=64, help="--impl reference-gpu: images per step (capped at the config's batch)")
    ap.add_argument("--ahead", type=int, default=1,
                    help="steps launched ahead of the one being finished (both loops).  Measured on B200: 1 / 2 / 3 ahead give 3.67 / 3.66 / "
                         "3.67 ms resident and 3.88 / 4.13 / 4.16 ms end to end (more uploads in flight), so 1 is the default")
    ap.add_argument("--streams", type=int, default=2,
                    help="CUDA streams that consecutive steps alternate over: 2 (default) = two batches overlap on the GPU, "
                         "1 = strictly one batch at a time")
    ap.add_argument("--gather-every", type=int, default=32,
                    help="N>1 only: steps a sweep of the detection exchange holds at most before it is drained (barrier + unpack)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.rotate <= 0:
        b = args.batch or WORKLOADS[args.config]["B"]
        args.rotate = max(2, int(math.ceil(140e6 / (b * 3 * 224 * 224 * 4))))
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    if args.impl == "reference":
        return run_reference(args, rank, world)
    if args.impl == "reference-gpu":
        return run_reference_gpu(args) if rank == 0 else 0
    return run_b200(args, rank, world, local_rank)


if __name__ == "__main__":
    sys.exit(main())
