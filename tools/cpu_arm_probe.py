"""CPU-arm probe (run on the GPU box): which thread count / pinning / batch size the reference's CPU path
is fastest with.  Prints the CPU topology and a table; decides nothing."""
import os, sys, time, subprocess
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import orc
import bench
print(subprocess.run("lscpu | egrep 'Model name|Socket|Core|Thread|NUMA|MHz'; lscpu -e=CPU,CORE,SOCKET,NODE | head -6; lscpu -e=CPU,CORE,SOCKET,NODE | sed -n '33,36p;65,68p;97,100p'",
                     shell=True, capture_output=True, text=True).stdout)
print("allowed cpus:", len(os.sched_getaffinity(0)))
mode, x = bench.cpu_streams_on_host("1200", 48000, 192000, 64)
kind = "reference-dfti" if orc.have_ref_dfti() else "reference"
orc.rx_many(mode, x[:1], nsamples=192000, nthreads=1, kind=kind)
def batch(n):
    return np.tile(x, ((n + 63) // 64, 1))[:n]
for n in (2048, 8192):
    b = batch(n)
    for nt in (32, 64, 128):
        ts = []
        for _ in range(3):
            t = time.perf_counter(); orc.rx_many(mode, b, nsamples=192000, nthreads=nt, kind=kind); ts.append(time.perf_counter() - t)
        print("rx_many   n=%5d nt=%3d            : %s Ms/s" % (n, nt, [round(n * 192000 / t / 1e6) for t in ts]), flush=True)
    for pin in ("cores", "compact", "none"):
        os.environ["ORC_POOL_PIN"] = pin
        for nt in (32, 64, 128):
            p = orc.RxPool(mode, nt, kind); p.load(b, 192000)
            ts = [p.run()[0] for _ in range(4)]
            p.close()
            print("pool      n=%5d nt=%3d pin=%-8s: %s Ms/s" % (n, nt, pin, [round(n * 192000 / t / 1e6) for t in ts]), flush=True)
