#!/bin/bash
# round-2 call A: the GPU suite, the full bench line (all BASELINE configs), a noise sweep of the headline,
# and the CPU reference arm twice (stability).
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv > gpurun_out/r2a_smi.txt; nproc >> gpurun_out/r2a_smi.txt; free -g >> gpurun_out/r2a_smi.txt
( time timeout 1200 python -m pytest tests -m gpu -x -q ) > gpurun_out/r2a_pytest.log 2>&1
tail -3 gpurun_out/r2a_pytest.log
( time timeout 1500 python bench.py --steps 5 --warmup 3 ) > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err
tail -c 600 gpurun_out/r2a_bench.err
for s in 0.2 0.35 0.5 0.7 1.0; do
  timeout 300 python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-configs --awgn $s 2>/dev/null | tail -1 | python -c "
import sys,json; d=json.loads(sys.stdin.read()); r=d['roofline']; print('awgn $s', round(d['value']), 'frac', round(r['frac'],3), 'cand/frame', round(r['candidates_per_frame'],2), r['decode_check'])" >> gpurun_out/r2a_awgn_sweep.txt
done
cat gpurun_out/r2a_awgn_sweep.txt
for i in 1 2; do timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2a_reference_$i.json 2>gpurun_out/r2a_reference_$i.err; done
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2a_bench.json').read().strip().splitlines()[-1])
print('value',round(d['value']),'frac',round(d['roofline']['frac'],3),'e2e',round(d['e2e']['value']),'s16',round(d['e2e']['s16_ingest']['value']), 'cpu', d['cpu_baseline']['value'], d['cpu_baseline']['seconds_all_passes'])
for c in d['configs']: print(c['key'], round(c['value']), round(c['roofline_frac'],3), round(c['candidates_per_frame'],2), c['decode_check'])
for i in (1,2):
    r=json.loads(open('gpurun_out/r2a_reference_%d.json'%i).read().strip().splitlines()[-1]); print('reference arm', i, round(r['value']), r['cpu_baseline']['seconds_all_passes'], r['cpu_baseline']['cores'])
PY
