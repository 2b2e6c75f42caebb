#!/bin/bash
# usage (on the GPU box): tools/ab_variants.sh name1 name2 ...  -> per-kernel forward times of each variant library
# (variants built beforehand with tools/build_variant.sh; "base" = the in-tree library)
for v in "$@"; do
  if [ "$v" = base ]; then unset VB200_LIB; else export VB200_LIB=$PWD/vampire_b200/_lib/variants/libvb200_$v.so; fi
  python bench.py --steps 10 $BENCH_ARGS --no-train-probe --no-aten-baseline --no-cpu-baseline --no-e2e --no-uncached 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
k=d['aux']['kernels']
print('%-14s step %.4f serial %.4f | '%('$v', d['ms_per_step'], d['aux']['ms_per_step_branches_serialised']) + ' '.join('%s %.4f'%(n,v['ms_per_step']) for n,v in k.items()))
"
done
