#!/bin/bash
mkdir -p gpurun_out
N=${1:-2}
( time timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 ) > gpurun_out/r2_bench_n$N.json 2> gpurun_out/r2_bench_n$N.err
tail -c 400 gpurun_out/r2_bench_n$N.err
python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/r2_bench_n$N.json').read().strip().splitlines() if l.startswith('{')][-1])
print('n_gpus', d['n_gpus'], 'value',round(d['value']),'frac',round(d['roofline']['frac'],3),'e2e',round(d['e2e']['value']),'s16',round(d['e2e']['s16_ingest']['value']))
for c in d['configs']: print(c['key'], round(c['value']), round(c['roofline_frac'],3), c['streams_per_gpu'], c['frames_decoded'])
PY
