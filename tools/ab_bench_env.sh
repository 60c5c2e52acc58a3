#!/bin/bash
# A/B of an environment switch on ONE box: tools/ab_bench_env.sh VAR v1 v2 ...  -> resident / e2e ms per step and the residual pass
# per launch for each value (bench.py without its baseline / variant legs).
var=$1; shift
for v in "$@"; do
  env $var=$v timeout 170 python bench.py --no-cpu-baseline --no-gpu-baseline --no-variants --sustain-s 0 --steps 40 2>/dev/null | tail -1 | \
    python -c "
import json,sys
d=json.loads(sys.stdin.read())
kb=d['kernel_breakdown']
print('$var=$v', 'resident %.3f ms' % d['ms_per_step'], 'e2e %.3f ms' % d['e2e']['ms_per_step'], 'add_layernorm768 %.1f us x%d' % (1e3*kb['add_layernorm768']['ms_per_step']/kb['add_layernorm768']['launches_per_step'], kb['add_layernorm768']['launches_per_step']), 'gemm2 share %.3f' % d['roofline']['share_of_step'], 'sum of kernels %.3f ms' % sum(k['ms_per_step'] for k in kb.values()))
"
done
