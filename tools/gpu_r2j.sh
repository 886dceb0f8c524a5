#!/bin/bash
mkdir -p gpurun_out
timeout 900 python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu --only-config cfg > gpurun_out/r2j_bench.json 2>gpurun_out/r2j_err.txt
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2j_bench.json').read().strip().splitlines()[-1])
print('value',round(d['value']),'frac',round(d['roofline']['frac'],3), 'cand/frame', round(d['roofline']['candidates_per_frame'],2))
for c in d['configs']: print('  ', c['key'], round(c['value']), round(c['roofline_frac'],3), round(c['candidates_per_frame'],2), c['decode_check'])
PY
bash tools/gpu_ncu.sh r2j_bell103 --mode 300 --streams 8192 --amplitude 0.5
