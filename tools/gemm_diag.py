"""Mainloop diagnostics for the CTA-pair GEMM: normal vs TMA-only vs MMA-only (HOIGEN_GEMM_DEBUG=0/1/2)."""
import os, subprocess, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if len(sys.argv) > 1:
    import torch
    from hoigen_b200 import _cabi
    dev = torch.device("cuda:0")
    shapes = [(215296, 256, 64, 2256), (16384, 1024, 256, 2256)] if os.environ.get('CONV_SHAPES') else [(12608, 3072, 768, 2256), (12608, 768, 3072, 2256), (12608, 2304, 768, 2256), (12608, 768, 768, 2256)]
    for (M, N, K, bn) in shapes:
        a = torch.randn(M, K, device=dev).bfloat16(); w = torch.randn(N, K, device=dev).bfloat16()
        o = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
        for _ in range(3): _cabi.gemm_bf16(a, w, out_bf16=o, block_n=bn)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10): _cabi.gemm_bf16(a, w, out_bf16=o, block_n=bn)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        print(f"debug={os.environ.get('HOIGEN_GEMM_DEBUG','0')} {M}x{N}x{K} bn={bn}: {ms*1e3:.1f} us  {2.0*M*N*K/ms/1e9:.0f} TF-equivalent", flush=True)
else:
    for d in ("0", "1", "3", "4"):
        subprocess.run([sys.executable, __file__, "run"], env=dict(os.environ, HOIGEN_GEMM_DEBUG=d))
