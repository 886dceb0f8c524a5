"""tools/write_traffic.py <metrics.txt from tools/ncu_summary.sh> <mode> <rate> <streams> <nsamples> [key]:
adds (or replaces) one capture in profiles/r2_traffic.json -- the DRAM bytes of one rx-kernel launch from an
`ncu --set full` report, tied to the kernel SOURCE it was taken on (bench.kernel_source_hash()).  bench.py
reports `roofline.traffic` from it only while that hash still matches."""
import json, os, re, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
metrics, mode, rate, streams, nsamples = sys.argv[1], sys.argv[2], int(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5])
vals = {}
for ln in open(metrics):
    m = re.match(r"(dram__bytes_(read|write)\.sum)\s+([0-9.]+)\s+(\w+)", ln)
    if m:
        scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[m.group(4)]
        vals[m.group(2)] = float(m.group(3)) * scale
    m = re.match(r"Kernel Name\s+(.*)", ln)
    if m:
        vals["kernel"] = m.group(1).strip()[:80]
p = os.path.join(ROOT, "profiles", "r2_traffic.json")
try:
    doc = json.load(open(p))
except Exception:
    doc = {"what": "DRAM bytes per rx-kernel launch from ncu --set full captures (dram__bytes_read.sum + dram__bytes_write.sum); "
                   "valid only for the kernel source whose hash is recorded with each capture", "captures": []}
w = {"mode": mode, "rate": rate, "streams": streams, "nsamples": nsamples}
doc["captures"] = [c for c in doc["captures"] if c["workload"] != w]
doc["captures"].append({"workload": w, "kernel": vals.get("kernel"), "dram_bytes_read": vals["read"], "dram_bytes_write": vals["write"],
                        "kernel_source_sha16": bench.kernel_source_hash(), "report": os.path.basename(metrics)})
json.dump(doc, open(p, "w"), indent=1)
print(json.dumps(doc["captures"][-1]))
