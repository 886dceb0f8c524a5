#!/bin/bash
# quick A/B of the prefix-table kernel: the BASELINE configurations with FSK_B200_PREFIX=1 (one step each)
mkdir -p gpurun_out
FSK_B200_PREFIX=${PFX:-1} timeout 600 python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu "$@" > gpurun_out/r2p_quick.json 2> gpurun_out/r2p_quick.err
tail -c 400 gpurun_out/r2p_quick.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2p_quick.json').read().strip().splitlines()[-1])
print('headline', round(d['value']), 'frac', round(d['roofline']['frac'],3), d['roofline']['decode_check'], d['roofline'].get('kernel_variant',''))
for c in d['configs']: print('  ', c['key'], round(c['value']), round(c['roofline_frac'],3), 'cand/frame', round(c['candidates_per_frame'],2), c['decode_check']['fraction_exact'], c.get('kernel','')[:60])
PY
