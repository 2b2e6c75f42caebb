#!/bin/bash
# usage (GPU box): tools/ab_env.sh "NAME=VALUE" ...  -> per-kernel forward times with each environment setting ("-" = none)
for v in "$@"; do
  if [ "$v" = "-" ]; then pre=""; else pre="env $v"; fi
  $pre python bench.py --steps 10 $BENCH_ARGS --no-train-probe --no-aten-baseline --no-cpu-baseline --no-e2e --no-uncached 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
k=d['aux']['kernels']
print('%-26s step %.4f serial %.4f | '%('$v', d['ms_per_step'], d['aux']['ms_per_step_branches_serialised']) + ' '.join('%s %.4f'%(n,v['ms_per_step']) for n,v in k.items()))
"
done
