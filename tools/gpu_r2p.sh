#!/bin/bash
# prefix-table search (mode 3) on hardware: parity of the variant and of the BASELINE configurations, then every
# configuration with the prefix search everywhere (FSK_B200_PREFIX=1), by default (auto) and without it (0)
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "variants or baseline_configs or reference_vectors or noise_sweep or s16_resident" ) > gpurun_out/r2p_pytest.log 2>&1; tail -3 gpurun_out/r2p_pytest.log
for P in 1 0; do
    FSK_B200_PREFIX=$P timeout 600 python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu > gpurun_out/r2p_bench_prefix$P.json 2> gpurun_out/r2p_bench_prefix$P.err
    tail -c 300 gpurun_out/r2p_bench_prefix$P.err
done
python - <<'PY'
import json
for P in "10":
    try:
        d=json.loads(open('gpurun_out/r2p_bench_prefix%s.json'%P).read().strip().splitlines()[-1])
    except Exception as e:
        print('prefix', P, 'failed', e); continue
    print('PREFIX=%s'%P, 'cfg2 clean', round(d['value']), 'frac', round(d['roofline']['frac'],3), d['roofline']['decode_check'])
    for c in d['configs']: print('  ', c['key'], round(c['value']), round(c['roofline_frac'],3), 'cand/frame', round(c['candidates_per_frame'],2), c['decode_check'], c.get('kernel',''))
PY
