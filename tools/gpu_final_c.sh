#!/bin/bash
# closing measurements of round 2 (prefix-table kernel in the defaults), one GPU:
#   the whole -m gpu suite, the default bench line, the reference arm, the launch list, and one ncu --set full
#   capture per BASELINE configuration, summarised ON THE BOX (the reports are ~20-30 MB each)
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/r2c_pytest.log 2>&1; grep -E "passed|failed" gpurun_out/r2c_pytest.log | tail -1
( time timeout 1500 python bench.py ) > gpurun_out/r2c_bench_n1.json 2> gpurun_out/r2c_bench_n1.err; tail -c 300 gpurun_out/r2c_bench_n1.err
timeout 600 python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/r2c_bench_reference_n1.json 2>/dev/null
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2c_launches_ncu.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e > gpurun_out/r2c_launches_bench.log 2>&1
cap() { tag=$1; shift
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_rx -s 1 -c 1 -f -o /tmp/$tag python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu --no-configs "$@" > gpurun_out/$tag.log 2>&1
  tools/ncu_summary.sh /tmp/$tag.ncu-rep /tmp/$tag > /dev/null 2>&1
  cp /tmp/$tag.metrics.txt gpurun_out/${tag}_metrics.txt
  python tools/ncu_hotspots.py /tmp/$tag.source.csv 40 > gpurun_out/${tag}_source_hotspots.txt
  grep -E "Kernel Name|time_duration|dram__bytes|issue_active|inst_executed.sum" gpurun_out/${tag}_metrics.txt | cut -c1-160
}
cap r2c_k_rx_ncu_cfg2
cap r2c_k_rx_ncu_bell103 --mode 300 --streams 32768 --amplitude 0.5
cap r2c_k_rx_ncu_rtty --mode rtty --rate 8000 --streams 262144 --nsamples 32000
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2c_bench_n1.json').read().strip().splitlines()[-1])
print('value',round(d['value']),'frac',round(d['roofline']['frac'],3),'e2e',round(d['e2e']['value']),'s16',round(d['e2e']['s16_ingest']['value']), 'cpu', round(d['cpu_baseline']['value']))
for c in d['configs']: print(c['key'], round(c['value']), round(c['roofline_frac'],3), round(c['candidates_per_frame'],2), c['decode_check']['fraction_exact'], c['kernel'][:48])
r=json.loads(open('gpurun_out/r2c_bench_reference_n1.json').read().strip().splitlines()[-1]); print('reference arm', round(r['value']))
PY
du -sh gpurun_out
