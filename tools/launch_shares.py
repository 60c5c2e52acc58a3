"""ncu launch list (--csv --log-file, gpu__time_duration.sum) -> markdown table of per-kernel launch counts, time and share."""
import csv
import re
import sys
from collections import defaultdict

src, dst, title = sys.argv[1], sys.argv[2], sys.argv[3]
rows = [r for r in csv.reader(l for l in open(src) if not l.startswith("==")) if len(r) > 5]
hdr = rows[0]
ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg = defaultdict(lambda: [0, 0.0])
for r in rows[1:]:
    v = float(r[iv].replace(",", ""))
    v = v / 1e3 if r[iu] in ("ns", "nsecond") else (v * 1e3 if r[iu] in ("ms", "msecond") else v)
    name = re.sub(r"\(.*", "", r[ik]).replace("hoigen::", "")[:40]
    agg[name][0] += 1
    agg[name][1] += v
tot = sum(v[1] for v in agg.values())
with open(dst, "w") as f:
    f.write(f"# {title}\n\n(raw list: {src.replace('gpurun_out', 'profiles')}; times are cold-cache and serialised — compare SHARES, not absolutes)\n\n")
    f.write(f"total {tot:.1f} us over {sum(v[0] for v in agg.values())} launches\n\n| kernel | launches | us | share |\n|---|---|---|---|\n")
    for k, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        f.write(f"| {k} | {n} | {us:.1f} | {us / tot:.3f} |\n")
