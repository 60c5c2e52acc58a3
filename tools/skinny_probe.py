"""Long-K / small-N GEMM shapes of the scoring head under every tile configuration (cold L2 and hot) — DESIGN.md 4."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from hoigen_b200 import _cabi
dev = torch.device("cuda:0")
_cabi.init(dev)
for (M, N, K) in [(7680, 117, 4096), (64, 117, 4096), (7680, 4096, 512), (64, 4096, 2048), (7680, 117, 512), (7680, 117, 1536)]:
    a = torch.randn(M, K, device=dev).bfloat16(); w = torch.randn(N, K, device=dev).bfloat16()
    ld = (N + 3) // 4 * 4
    o = torch.zeros(M, ld, device=dev)
    ob = torch.zeros(M, (N + 7) // 8 * 8, device=dev, dtype=torch.bfloat16)
    big = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    for bn in (0, 64, 128, 256, 2128, 2192, 2256):
        if bn in (2128, 2192, 2256) and M <= 128: continue
        if bn % 1000 > 64 and N <= (bn % 1000) // 2: continue
        kw = dict(out_f32=o[:, :N], residual=o[:, :N]) if N == 117 else dict(out_bf16=ob[:, :N])
        try:
            for _ in range(3): _cabi.gemm_bf16(a, w, block_n=bn, **kw)
        except Exception as e:
            print(f"{M}x{N}x{K} bn={bn}: {e}"); continue
        torch.cuda.synchronize()
        ts = []
        for _ in range(8):
            big.zero_()                                   # flush L2 so A streams from HBM as in the pipeline... (phi is freshly written there)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); _cabi.gemm_bf16(a, w, block_n=bn, **kw); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        hot = []
        for _ in range(8):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); _cabi.gemm_bf16(a, w, block_n=bn, **kw); e1.record(); torch.cuda.synchronize()
            hot.append(e0.elapsed_time(e1))
        print(f"{M}x{N}x{K} bn={bn:5d}: cold {sorted(ts)[len(ts)//2]*1e3:7.1f} us   hot {sorted(hot)[len(hot)//2]*1e3:7.1f} us", flush=True)
