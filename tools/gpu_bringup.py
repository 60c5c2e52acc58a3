"""GPU bring-up checks run under gpurun (each case in its own process so a trap does not poison the rest).

    python tools/gpu_bringup.py            # runs every case in a subprocess with a timeout
    python tools/gpu_bringup.py --case gemm_basic
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def _gemm_ref(a, w, bias, act, colscale, residual):
    import torch
    v = a.float() @ w.float().t()
    if bias is not None:
        v = v + bias
    if act == 1:
        v = v * torch.sigmoid(1.702 * v)
    elif act == 2:
        v = torch.relu(v)
    if colscale is not None:
        v = v * colscale
    if residual is not None:
        v = v + residual
    return v


def case_gemm_basic():
    import torch
    from hoigen_b200 import _cabi
    torch.manual_seed(0)
    dev = torch.device("cuda:0")
    out = {}
    for (M, N, K, bn) in [(128, 64, 64, 64), (128, 128, 64, 128), (128, 256, 64, 256), (128, 256, 256, 256),
                          (256, 256, 768, 256), (300, 200, 136, 0), (12608, 768, 768, 0), (12608, 2304, 768, 0)]:
        a = torch.randn(M, K, device=dev).bfloat16()
        w = (torch.randn(N, K, device=dev) / K ** 0.5).bfloat16()
        o = torch.full((M, N), float("nan"), device=dev)
        _cabi.gemm_bf16(a, w, out_f32=o, block_n=bn)
        torch.cuda.synchronize()
        ref = _gemm_ref(a, w, None, 0, None, None)
        err = (o - ref).abs().max().item()
        o2 = torch.empty_like(o)
        _cabi.gemm_bf16(a, w, out_f32=o2, simt=True)
        torch.cuda.synchronize()
        err_simt = (o2 - ref).abs().max().item()
        out[f"{M}x{N}x{K}_bn{bn}"] = {"max_abs_err": err, "simt_err": err_simt, "ref_absmax": ref.abs().max().item()}
        print(f"gemm {M}x{N}x{K} bn={bn}: err={err:.3e} simt_err={err_simt:.3e}", flush=True)
    return out


def case_gemm_epilogue():
    import torch
    from hoigen_b200 import _cabi
    torch.manual_seed(1)
    dev = torch.device("cuda:0")
    out = {}
    M, N, K = 1000, 776, 320
    a = torch.randn(M, K, device=dev).bfloat16()
    w = (torch.randn(N, K, device=dev) / K ** 0.5).bfloat16()
    bias = torch.randn(N, device=dev)
    cs = torch.rand(N, device=dev) + 0.5
    res = torch.randn(M, N, device=dev)
    for act in (0, 1, 2):
        for bn in (64, 128, 256):
            of = res.clone()
            ob = torch.zeros(M, N, device=dev, dtype=torch.bfloat16)
            _cabi.gemm_bf16(a, w, bias=bias, colscale=cs, act=act, residual=of, out_f32=of, out_bf16=ob, block_n=bn)
            torch.cuda.synchronize()
            ref = _gemm_ref(a, w, bias, act, cs, res)
            e1 = (of - ref).abs().max().item()
            e2 = (ob.float() - ref).abs().max().item()
            out[f"act{act}_bn{bn}"] = {"f32_err": e1, "bf16_err": e2}
            print(f"epilogue act={act} bn={bn}: f32 err={e1:.3e} bf16 err={e2:.3e}", flush=True)
    # odd N (scalar tail path), bf16 only
    N2 = 117
    w2 = (torch.randn(N2, K, device=dev) / K ** 0.5).bfloat16()
    o = torch.zeros(M, 120, device=dev)
    _cabi.gemm_bf16(a, w2, out_f32=o[:, :N2])
    torch.cuda.synchronize()
    ref = _gemm_ref(a, w2, None, 0, None, None)
    out["n117"] = {"err": (o[:, :N2] - ref).abs().max().item(), "pad_untouched": bool((o[:, N2:] == 0).all().item())}
    print("n117", out["n117"], flush=True)
    return out


def case_gemm_perf():
    import torch
    from hoigen_b200 import _cabi
    dev = torch.device("cuda:0")
    out = {}
    for (M, N, K) in [(12608, 2304, 768), (12608, 768, 768), (12608, 3072, 768), (12608, 768, 3072),
                      (12608, 768, 64), (7680, 117, 4096), (7680, 4096, 512), (12608, 512, 768), (8192, 8192, 8192)]:
        a = torch.randn(M, K, device=dev).bfloat16()
        w = torch.randn(N, K, device=dev).bfloat16()
        o = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
        for bn in (0, 256, 2256, 2192, 2128):
            if bn and N < (bn % 2000) // 2:
                continue
            for _ in range(3):
                _cabi.gemm_bf16(a, w, out_bf16=o, block_n=bn)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            iters = 20
            e0.record()
            for _ in range(iters):
                _cabi.gemm_bf16(a, w, out_bf16=o, block_n=bn)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / iters
            tf = 2.0 * M * N * K / ms / 1e9
            out[f"{M}x{N}x{K}_bn{bn}"] = {"ms": ms, "tflops": tf}
            print(f"perf {M}x{N}x{K} bn={bn}: {ms*1e3:.1f} us  {tf:.1f} TFLOP/s", flush=True)
        # cuBLAS for context
        wt = w.t().contiguous()
        for _ in range(3):
            torch.matmul(a, wt)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            torch.matmul(a, wt)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 20
        out[f"{M}x{N}x{K}_cublas"] = {"ms": ms, "tflops": 2.0 * M * N * K / ms / 1e9}
        print(f"perf {M}x{N}x{K} cublas: {ms*1e3:.1f} us  {2.0*M*N*K/ms/1e9:.1f} TFLOP/s", flush=True)
    return out


CASES = {k[len("case_"):]: v for k, v in list(globals().items()) if k.startswith("case_")}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--case", default=None)
    ap.add_argument("--only", default=None, help="comma separated subset of cases for the driver mode")
    ap.add_argument("--timeout", type=int, default=240)
    args = ap.parse_args()
    os.makedirs("gpurun_out", exist_ok=True)
    if args.case:
        res = CASES[args.case]()
        with open(f"gpurun_out/bringup_{args.case}.json", "w") as f:
            json.dump(res, f, indent=1)
        return 0
    names = args.only.split(",") if args.only else list(CASES)
    rc_all = 0
    for name in names:
        t0 = time.time()
        try:
            r = subprocess.run([sys.executable, __file__, "--case", name], timeout=args.timeout)
            rc = r.returncode
        except subprocess.TimeoutExpired:
            rc = -999
        print(f"=== case {name}: rc={rc} ({time.time()-t0:.1f}s)", flush=True)
        rc_all = rc_all or rc
    return rc_all


if __name__ == "__main__":
    sys.exit(main())
