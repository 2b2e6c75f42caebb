#!/bin/bash
# usage: tools/bench_summary.sh [bench args]   -- one-line per-kernel summary of a bench run
python bench.py "$@" 2>&1 | grep '^{' | python -c "
import json,sys
for l in sys.stdin:
    d=json.loads(l)
    print('ms/step', round(d['ms_per_step'],3), 'serial', round(d['aux'].get('ms_per_step_branches_serialised',0),3), {k:(round(v['ms_per_step'],3), round(v.get('frac_of_peak',0),3)) for k,v in d['aux']['kernels'].items()})
"
