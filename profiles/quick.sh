#!/bin/bash
# quick GPU check: GPU tests + one bench line (no profiler)
TAG=${1:-q}
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" || exit 1
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -15
timeout 600 python bench.py ${@:2} > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
python - <<PY
import json
d=json.load(open('gpurun_out/bench_$TAG.json'))
print('value %.4g triples/s  ms/step %.4f  frac %.3f  e2e %.4g' % (d['value'], d['ms_per_step'], d['roofline']['frac'], d['e2e']['value']))
for s in d.get('sweep', []): print(s)
if 'score_topk' in d: print('score', d['score_topk']['value'], d['score_topk']['ms_per_step'], d['score_topk']['roofline']['achieved'])
PY
tail -3 gpurun_out/bench_$TAG.err
