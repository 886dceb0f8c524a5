#!/bin/bash
mkdir -p gpurun_out
run() { label=$1; shift; timeout 300 python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu --no-configs "$@" 2>/dev/null | tail -1 | python -c "
import sys,json; d=json.loads(sys.stdin.read()); r=d['roofline']; print('$label', round(d['value']), 'frac', round(r['frac'],3), 'cand/frame', round(r['candidates_per_frame'],2))"; }
run "rtty G=auto" --mode rtty --rate 8000 --streams 262144 --nsamples 32000
run "rtty G=8" --mode rtty --rate 8000 --streams 262144 --nsamples 32000 --lanes 8
run "rtty G=32" --mode rtty --rate 8000 --streams 262144 --nsamples 32000 --lanes 32
run "bell103 G=auto" --mode 300 --streams 32768 --amplitude 0.5
run "bell103 G=8" --mode 300 --streams 32768 --amplitude 0.5 --lanes 8
run "bell103 G=32" --mode 300 --streams 32768 --amplitude 0.5 --lanes 32
FSK_B200_MULTI=0 run "bell103 per-candidate" --mode 300 --streams 32768 --amplitude 0.5
FSK_B200_MULTI=1 run "bell103 hybrid" --mode 300 --streams 32768 --amplitude 0.5
FSK_B200_MULTI=2 run "cfg2 multi-always awgn" --awgn 0.35
FSK_B200_MULTI=1 run "cfg2 hybrid awgn" --awgn 0.35
FSK_B200_MULTI=1 run "cfg2 hybrid clean"
run "same" --mode same --streams 131072 --nsamples 24000
run "same G=16" --mode same --streams 131072 --nsamples 24000 --lanes 16
