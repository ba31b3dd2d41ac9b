#!/bin/bash
# usage: tools/sweep.sh <tag> <variant...>   -> gpurun_out/<tag>_sweep.jsonl  (one bench line per variant)
tag=$1; shift
mkdir -p gpurun_out
: > gpurun_out/${tag}_sweep.jsonl
for spec in "$@"; do
  v=${spec%%:*}; mb=0; [ "$spec" != "$v" ] && mb=${spec##*:}
  export RBX_L2_PERSIST_MB=$mb
  RBX_LIB_PATH=$PWD/build/variants/$v.so python bench.py --steps 30 --warmup 5 --no-cpu-baseline $BENCH_ARGS 2>>gpurun_out/${tag}_sweep.err | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); d['variant'] = '$spec'; print(json.dumps(d))" >> gpurun_out/${tag}_sweep.jsonl
done
python - <<PY
import json
for l in open('gpurun_out/${tag}_sweep.jsonl'):
    d = json.loads(l); k = d['kernels']
    print('%-16s value %6.1fM  e2e %6.1fM  ' % (d['variant'], d['value']/1e6, d['e2e']['value']/1e6) + '  '.join('%s %.1fus' % (n.split('(')[0][-12:], v['ms']*1e3) for n, v in k.items()))
PY
