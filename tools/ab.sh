#!/bin/bash
# A/B two builds of the library on the same box: tools/ab.sh <stage> ; expects hair-gs_b200/lib/variant_{a,b}.so
for r in 1 2; do for v in a b; do
  HGS_LIBRARY=$PWD/hair-gs_b200/lib/variant_$v.so python bench.py --steps 50 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('variant $v', d['value'], d['e2e']['value'], {k: v['ms_per_launch'] for k, v in d['stages'].items() if k in '$1'.split(',')})"
done; done
