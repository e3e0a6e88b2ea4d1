"""bench.py's reference arm runs on CPU (the reference's own C++ train(), oracle/_ref, or the oracle port): check that it prints ONE JSON line
with the keys the driver reads, on the metric / config of BASELINE.json."""
import json
import os
import subprocess
import sys

from conftest import ROOT


def test_reference_arm_prints_the_contract_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0", "--ref-sample-tokens", "200000"], capture_output=True,
                       text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [x for x in r.stdout.strip().splitlines() if x.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    base = json.load(open(os.path.join(ROOT, "BASELINE.json")))
    assert d["impl"] == "reference" and d["higher_is_better"] is True and d["unit"] == "tokens/s"
    assert d["metric"].split("(")[0].strip() == base["metric"].split("(")[0].strip()
    for k in ("value", "n_gpus", "steps", "warmup", "ms_per_step", "scaling", "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] == 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["value"] > 1e4  # tokens/s of a single CPU core: hundreds of thousands


def _record(name):
    with open(os.path.join(ROOT, "profiles", name)) as f:
        return json.loads([x for x in f.read().strip().splitlines() if x.startswith("{")][-1])


def test_committed_bench_record_satisfies_the_contract():
    """profiles/r02h_bench.json is what `python bench.py` printed on a B200 at the end of round 2: every key the driver reads, and numbers that
    follow from one another (value = tokens / time, frac = achieved / peak, achieved = algorithmic bytes / kernel time)."""
    d = _record("r02h_bench.json")
    base = json.load(open(os.path.join(ROOT, "BASELINE.json")))
    assert d["metric"].split("(")[0].strip() == base["metric"].split("(")[0].strip() and d["unit"] == "tokens/s" and d["higher_is_better"] is True
    for k in ("value", "n_gpus", "steps", "warmup", "ms_per_step", "scaling", "vs_baseline", "dtype", "data", "config", "roofline", "cpu_baseline", "e2e", "clocks", "gpu_launches"):
        assert k in d, k
    assert d["n_gpus"] == 1 and d["warmup"] >= 3 and d["scaling"] == "weak" and d["vs_baseline"] is None and d["data"] == "synthetic" and "workload" in d["config"]
    tokens = 100000000
    assert abs(d["value"] - tokens / (d["ms_per_step"] / 1e3)) / d["value"] < 1e-6
    r = d["roofline"]
    assert r["bound"] == "hbm" and r["unit"] == "GB/s" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    assert abs(r["achieved"] - r["alg_bytes_per_step"] / (r["kernel_ms_per_step"] / 1e3) / 1e9) / r["achieved"] < 1e-6
    assert 0 < r["frac"] < 1 and 0 < r["kernel_share_of_step"] <= 1
    e = d["e2e"]
    assert e["unit"] == "tokens/s" and e["h2d_bytes_per_step"] == d["config"]["corpus_bytes"] and e["d2h_bytes_per_step"] > 0
    assert e["value"] < d["value"]  # copies inside the timed region: end to end cannot beat the device-resident number
    assert abs(e["value"] - tokens / (e["ms_per_step"] / 1e3)) / e["value"] < 1e-6
    c = d["cpu_baseline"]
    assert c["kind"] == "reference" and c["cores"] == 1 and c["unit"] == "tokens/s" and "sample" in c and 0 < c["value"] < d["value"]
    assert d["clocks"]["sm_mhz"] > 0.9 * d["clocks"]["sm_max_mhz"] and not set(d["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    assert d["gpu_launches"] > 0
    p = d["parity"]
    assert p["digest_ok"] and p["checksum_ok"] and p["patterns_ok"] and p["passes_ok"]  # the timed model IS the reference's model


def test_committed_traffic_capture_matches_the_kernel_sources():
    """profiles/traffic.json names the sha of the kernel sources it was captured from; bench.py prints roofline.traffic only while it matches."""
    import importlib.util

    spec = importlib.util.spec_from_file_location("bench", os.path.join(ROOT, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    t, why = bench.profile_traffic()
    assert t is not None, why
    assert abs(t["dram_bytes_per_step"] - sum((l["dram_read_gb"] + l["dram_write_gb"]) * 1e9 for l in t["launches"])) < 1e6
    alg = _record("r02h_bench.json")["roofline"]["alg_bytes_per_step"]
    assert 1.0 < t["dram_bytes_per_step"] / alg < 1.3  # DRAM traffic within 1.3x of the algorithmic bytes


def test_committed_multi_gpu_records_carry_the_parity_step():
    for n, name in ((2, "r02h_bench_n2.json"), (4, "r02h_bench_n4.json"), (8, "r02h_bench_n8.json")):
        d = _record(name)
        assert d["n_gpus"] == n and d["scaling"] == "weak" and d["parity"]["checksum_ok"] is True
        assert abs(d["value"] - n * 100000000 / (d["ms_per_step"] / 1e3)) / d["value"] < 1e-6  # whole-job tokens / max-over-ranks time
        assert d["e2e"]["value"] < d["value"]
