"""The reference arm of bench.py (CPU only): one JSON line with the contract's keys."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_json_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "0", "--cpu-seconds", "0.5"], stdout=subprocess.PIPE, stderr=subprocess.PIPE,
                         text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, lines
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "loci/s" and d["higher_is_better"] is True
    assert d["metric"] == "loci/sec for ols_iter (f64)" and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"] == {"value": d["value"], "unit": "loci/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"]
    # the reference arm never maps the product library (its inputs come from libpoolgen_synth.so) and reports the
    # "tight" CPU variant next to the faithful one
    assert d["maps_product_library"] is False
    assert d["cpu_baseline"]["tight"]["value"] > d["value"]


def test_traffic_is_read_from_the_committed_capture():
    """roofline.traffic comes from profiles/ncu_raw_r*.csv at run time (dram__bytes_read + dram__bytes_write of the
    streaming kernel and the fix-up kernel), not from a constant"""
    sys.path.insert(0, ROOT)
    import bench
    per_locus, src = bench.traffic_from_profile(1000, 4, 3)
    assert src and src.startswith("profiles/ncu_raw_r")
    assert 32288 <= per_locus <= 1.2 * 32288          # at least the algorithmic bytes, no wasted re-reads
    assert bench.traffic_from_profile(7, 7, 7) == (None, None)


def test_committed_bench_lines_follow_the_contract():
    """profiles/bench_*_r2.json (what bench.py printed on the GPU boxes): one JSON line each with the driver's keys, the
    roofline / cpu_baseline / e2e objects and the per-config summaries at the top level"""
    import glob
    import json
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    files = sorted(glob.glob(os.path.join(root, "profiles", "bench_*gpu_r2.json")))
    assert len(files) >= 3
    for f in files:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                    "vs_baseline", "dtype", "data", "config", "roofline", "e2e", "gpu_launches", "clocks", "c4"):
            assert key in d, (f, key)
        assert d["unit"] == "loci/s" and d["higher_is_better"] is True and d["scaling"] == "weak" and d["dtype"] == "f64"
        assert d["warmup"] >= 3 and d["gpu_launches"] > 0 and "workload" in d["config"] and "model" not in d["config"]
        r = d["roofline"]
        assert r["bound"] == "hbm" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9 and r["traffic"] > 0
        assert r["traffic_source"].startswith("profiles/ncu_raw_r")
        e = d["e2e"]
        assert e["unit"] == "loci/s" and e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0 and e["value"] < d["value"]
        assert not set(d["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
        if d["n_gpus"] == 1:
            cb = d["cpu_baseline"]
            assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] > 0 and "tight" in cb
            for key in ("c2", "c5", "text", "nelder_mead"):
                assert key in d, (f, key)
