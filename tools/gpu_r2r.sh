#!/bin/bash
run() { label=$1; shift; timeout 300 python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu --no-configs "$@" 2>/dev/null | tail -1 | python -c "
import sys,json; d=json.loads(sys.stdin.read()); r=d['roofline']; print('$label', round(d['value']), 'frac', round(r['frac'],3))"; }
for w in 1 2 3 4; do
run "bell103 wpb=$w" --mode 300 --streams 32768 --amplitude 0.5 --wpb $w
run "rtty wpb=$w" --mode rtty --rate 8000 --streams 262144 --nsamples 32000 --wpb $w
run "same wpb=$w" --mode same --streams 131072 --nsamples 24000 --wpb $w
run "cfg2awgn wpb=$w" --awgn 0.35 --wpb $w
done
