#!/bin/bash
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests -m gpu -x -q ) > gpurun_out/r2c_pytest.log 2>&1
tail -3 gpurun_out/r2c_pytest.log
FSK_B200_MULTI=0 timeout 600 python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu --no-configs > gpurun_out/r2c_bench_multi0.json 2>gpurun_out/r2c_err0.txt
timeout 900 python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu > gpurun_out/r2c_bench_multi1.json 2>gpurun_out/r2c_err1.txt
python - <<'PY'
import json
for f in ('gpurun_out/r2c_bench_multi0.json','gpurun_out/r2c_bench_multi1.json'):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as e:
        print(f, 'FAILED', e); print(open(f.replace('bench_multi','err').replace('.json','.txt')).read()[-1500:]); continue
    print(f, 'value',round(d['value']),'frac',round(d['roofline']['frac'],3), 'cand/frame', round(d['roofline']['candidates_per_frame'],2))
    for c in d['configs']: print('  ', c['key'], round(c['value']), round(c['roofline_frac'],3), round(c['candidates_per_frame'],2), c['decode_check'])
PY
