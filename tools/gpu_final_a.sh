#!/bin/bash
# final measurements, call A (1 GPU): suite, the default bench line, the reference arm twice, drop-in timing, launch list, ncu of cfg2
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/r2_final_pytest.log 2>&1; tail -3 gpurun_out/r2_final_pytest.log
( time timeout 1500 python bench.py ) > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err; tail -c 300 gpurun_out/r2_bench_n1.err
for i in a b; do timeout 600 python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/r2_bench_reference_n1_$i.json 2>/dev/null; done
python tools/dropin_timing.py > gpurun_out/r2_dropin_timing.txt 2>&1; tail -1 gpurun_out/r2_dropin_timing.txt
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches_ncu.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-configs > gpurun_out/r2_launches_bench.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_rx -s 1 -c 1 -f -o gpurun_out/r2_ncu_cfg2 python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu --no-configs > gpurun_out/r2_ncu_cfg2.log 2>&1
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_bench_n1.json').read().strip().splitlines()[-1])
print('value',round(d['value']),'frac',round(d['roofline']['frac'],3),'e2e',round(d['e2e']['value']),'s16',round(d['e2e']['s16_ingest']['value']), 'cpu', round(d['cpu_baseline']['value']), d['cpu_baseline']['seconds_all_passes'])
for c in d['configs']: print(c['key'], round(c['value']), round(c['roofline_frac'],3), round(c['candidates_per_frame'],2), c['decode_check'])
for i in 'ab':
    r=json.loads(open('gpurun_out/r2_bench_reference_n1_%s.json'%i).read().strip().splitlines()[-1]); print('reference arm', i, round(r['value']), 'best', round(r['cpu_baseline']['best']), r['cpu_baseline']['seconds_all_passes'], r['cpu_baseline']['cores'])
PY
du -sh gpurun_out
