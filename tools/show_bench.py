"""Pretty-print a bench.py JSON line (kernel breakdown) — local helper."""
import json
import sys

d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(f"value {d['value']:.1f} {d['unit']}  ms/step {d['ms_per_step']:.3f}  e2e {d['e2e']['value']:.1f} ({d['e2e']['ms_per_step']:.3f} ms)  "
      f"launches {d['gpu_launches']}  clocks {d['clocks']}")
r = d["roofline"]
print(f"roofline: achieved {r['achieved']:.1f} {r['unit']} / peak {r['peak']} = {r['frac']:.3f}; gemm share {r['share_of_step']:.3f}; "
      f"encoder-effective {d['encoder_tflops_effective']:.1f} TF")
tot = 0.0
for k, v in sorted(d["kernel_breakdown"].items(), key=lambda kv: -kv[1]["ms_per_step"]):
    tot += v["ms_per_step"]
    tf = f"{v['tflops']:.1f}TF" if v["tflops"] else ""
    gb = f"{v['gbs']:.0f}GB/s" if v["gbs"] else ""
    print(f"  {k:26s} n={v['launches_per_step']:5.1f} ms={v['ms_per_step']:.4f} us/launch={v['ms_per_step']/v['launches_per_step']*1e3:7.1f} {tf:>9s} {gb:>10s}")
print(f"  sum of kernels {tot:.3f} ms")
if d.get("cpu_baseline"):
    print("cpu_baseline", d["cpu_baseline"]["value"], d["cpu_baseline"]["cores"])
