"""bench.py's output contract, as far as it can be checked without a GPU: the reference arm
(`--impl reference`, the unmodified src/fsk.c on the host cores) prints ONE JSON line with the keys
the driver reads; under a multi-rank launch only rank 0 prints."""
import json
import os
import subprocess
import sys

import pytest

import orc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run_bench(extra_env=None):
    env = dict(os.environ, **(extra_env or {}))
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                        "--warmup", "0", "--cpu-streams", "8", "--nsamples", "48000"],
                       cwd=ROOT, env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=900)
    assert r.returncode == 0, r.stderr[-800:]
    return [ln for ln in r.stdout.decode().splitlines() if ln.strip()]


def test_reference_arm_prints_one_contract_line():
    lines = run_bench()
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference"
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better",
                "scaling", "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e", "gpu_launches"):
        assert key in d, key
    assert d["metric"].startswith("audio Msamples/s demodulated") and d["unit"] == "Msamples/s"
    assert d["higher_is_better"] is True and d["vs_baseline"] is None and d["gpu_launches"] == 0
    assert "workload" in d["config"] and "model" not in d["config"]
    cb = d["cpu_baseline"]
    assert cb["kind"] == ("reference" if orc.have_ref() else "port") and cb["cores"] >= 1 and cb["value"] > 0
    assert cb["value"] == d["value"] == d["e2e"]["value"]
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    if orc.have_ref():
        assert "not FFTW" in cb["sample"]           # the stand-in FFT is named, whichever it is


def test_reference_arm_other_ranks_stay_silent():
    assert run_bench({"RANK": "1", "LOCAL_RANK": "1", "WORLD_SIZE": "2"}) == []


def test_recorded_gpu_line_has_the_contract_keys():
    """profiles/r1_bench_n1.json: the line `python bench.py` printed on a B200 at the end of round 1."""
    d = json.loads(open(os.path.join(ROOT, "profiles", "r1_bench_n1.json")).read().strip().splitlines()[-1])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "roofline", "cpu_baseline", "e2e", "gpu_launches", "clocks"):
        assert key in d, key
    r = d["roofline"]
    assert r["bound"] == "hbm" and r["unit"] == "GB/s" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    assert r["traffic"] and 0.99 < r["traffic"] / r["algorithmic_bytes"] < 1.01      # every sample read once
    assert d["n_gpus"] == 1 and d["steps"] >= 1 and d["warmup"] >= 3 and d["gpu_launches"] >= d["steps"]
    assert d["e2e"]["h2d_bytes_per_step"] > 0 and d["e2e"]["d2h_bytes_per_step"] > 0
    assert d["e2e"]["value"] < d["value"]                       # host buffers: PCIe-bound, not a copy of `value`
    assert not set(d["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    assert "workload" in d["config"] and d["dtype"] == "f32" and d["data"] == "synthetic"


def test_recorded_round2_line_carries_every_baseline_configuration():
    """profiles/r2_bench_n1.json: the closing line of round 2 -- the headline with its ncu-backed traffic figure, and
    the `configs` block with every other BASELINE configuration, each with kernel time, roofline fraction, the
    kernel variant that ran and a decode check on >= 1 % of the streams."""
    d = json.loads(open(os.path.join(ROOT, "profiles", "r2_bench_n1.json")).read().strip().splitlines()[-1])
    r = d["roofline"]
    assert abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9 and r["frac"] >= 0.70          # the north star's bar
    assert r["traffic"] and 0.99 < r["traffic"] / r["algorithmic_bytes"] < 1.01
    assert not set(d["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    keys = [c["key"] for c in d["configs"]]
    for want in ("cfg2_1200_awgn", "cfg3_rtty_8k", "cfg4_bell103_offset0.00", "cfg4_bell103_awgn0.50", "cfg5_same_per_gpu"):
        assert any(k.startswith(want) for k in keys), (want, keys)
    for c in d["configs"]:
        assert c["value"] > 0 and 0 < c["roofline_frac"] < 1 and c["kernel_ms"] > 0 and c["kernel"].startswith("k_rx<")
        assert c["decode_check"]["streams"] * 100 >= c["streams_per_gpu"]
        if "awgn0.50" not in c["key"]:
            assert c["decode_check"]["fraction_exact"] == 1.0, c["key"]
    by = {c["key"]: c for c in d["configs"]}
    assert "prefix-table" in by["cfg3_rtty_8k_clean"]["kernel"] and "prefix-table" in by["cfg4_bell103_offset0.00"]["kernel"]
