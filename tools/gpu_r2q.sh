#!/bin/bash
run() { label=$1; shift; timeout 300 python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu --no-configs "$@" 2>/dev/null | tail -1 | python -c "
import sys,json; d=json.loads(sys.stdin.read()); r=d['roofline']; print('$label', round(d['value']), 'frac', round(r['frac'],3), 'cand/frame', round(r['candidates_per_frame'],2))"; }
run "clean slide"
FSK_B200_NO_SLIDE=1 run "clean noslide"
run "awgn slide" --awgn 0.35
FSK_B200_NO_SLIDE=1 run "awgn noslide" --awgn 0.35
run "awgn0.5 slide" --awgn 0.5
FSK_B200_NO_SLIDE=1 run "awgn0.5 noslide" --awgn 0.5
run "same slide" --mode same --streams 131072 --nsamples 24000
FSK_B200_NO_SLIDE=1 run "same noslide" --mode same --streams 131072 --nsamples 24000
