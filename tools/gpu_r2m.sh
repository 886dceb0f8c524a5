#!/bin/bash
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests -m gpu -x -q ) > gpurun_out/r2m_pytest.log 2>&1
tail -3 gpurun_out/r2m_pytest.log
FSK_B200_TRACE=1 timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu --no-configs > gpurun_out/r2m_bench.json 2>gpurun_out/r2m_err.txt
grep "fsk_b200 trace" gpurun_out/r2m_err.txt | tail -4
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2m_bench.json').read().strip().splitlines()[-1])
print('value',round(d['value']),'frac',round(d['roofline']['frac'],3)); print(json.dumps(d['e2e'])[:900])
PY
# device-resident int16 vs float: kernel time at the headline size
python - <<'PY'
import torch, time, sys
sys.path.insert(0,'.')
import minimodem_b200 as mm, bench
a = bench.parse.__wrapped__() if hasattr(bench.parse,'__wrapped__') else None
dev=torch.device('cuda:0')
wl = bench.Workload(mm, torch, dev, 0, "1200", 48000, 32768, 192000, 1.0)
eng = mm.RxEngine(wl.params)
x16 = torch.empty((wl.S, (wl.n+7)&~7), dtype=torch.int16, device=dev)
x16.zero_()
rows=1024
for s0 in range(0, wl.S, rows):
    x16[s0:s0+rows, :wl.stride] = (wl.x[s0:s0+rows]*32767.0).round().to(torch.int16)
mf = eng.max_frames(wl.n)
fr = torch.empty((wl.S, mf, 5), dtype=torch.int32, device=dev); st = torch.zeros((wl.S, mm.STATE_WORDS), dtype=torch.int32, device=dev)
for name, src in (("f32", wl.x), ("s16", x16)):
    for _ in range(3):
        st.zero_(); eng.rx_batch(src, nsamples=wl.n, max_frames=mf, frames=fr, states=st)
    torch.cuda.synchronize()
    e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    ts=[]
    for _ in range(3):
        st.zero_(); e0.record(); eng.rx_batch(src, nsamples=wl.n, max_frames=mf, frames=fr, states=st); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    ms=sum(ts)/3
    print(name, eng.last_kernel().split('>')[0], "%.2f ms"%ms, "%.0f Msamples/s"%(wl.S*wl.n/ms/1e3), "frames", int(mm.states_to_numpy(st)["nframes"].sum()))
PY
