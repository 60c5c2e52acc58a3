#!/bin/bash
# A/B of bench.py argument sets on ONE box: tools/ab_bench_args.sh "--ahead 1" "--ahead 2" ...  (bench.py without its baseline /
# variant legs; resident and end-to-end ms per step).
for a in "$@"; do
  timeout 170 python bench.py --no-cpu-baseline --no-gpu-baseline --no-variants --sustain-s 0 --steps 40 $a 2>/dev/null | tail -1 | \
    python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('$a', '| resident %.3f ms' % d['ms_per_step'], '| e2e %.3f ms' % d['e2e']['ms_per_step'])
"
done
