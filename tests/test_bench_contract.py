"""CPU check of bench.py's contract for the reference arm (the GPU arm needs a B200): one JSON line with the
keys the driver reads, produced by the CPU port of the reference on the FULL ensemble of the workload (the reference's
index! phase is table-parallel only, so a sample of a few tables would starve the host cores), for exactly the
steps / warm-up the driver asks for, with the same `config` dict the GPU arm emits."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_json_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "2", "--warmup", "1"],
                         capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.strip().splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "embedding_lookups_per_sec" and d["unit"] == "lookups/s"
    assert d["higher_is_better"] is True and d["value"] > 0 and d["ms_per_step"] > 0 and d["vs_baseline"] is None
    assert d["steps"] == 2 and d["warmup"] == 1                      # what the driver passed, not a clamp
    sys.path.insert(0, ROOT)
    import bench
    assert d["config"] == bench.common_config(1, "uniform")          # the dict the GPU arm emits too
    assert d["config"]["workload"].startswith("C2") and d["config"]["tables"] == 26
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] == len(os.sched_getaffinity(0)) and cb["value"] == d["value"]
    assert cb["sample"].startswith("all 26 tables") and "pinned" in cb["sample"]   # the full ensemble, not a sample of it
    assert d["e2e"] == {"value": d["value"], "unit": "lookups/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_nonzero_ranks_do_nothing():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"],
                         capture_output=True, text=True, timeout=60, cwd=ROOT, env=dict(os.environ, RANK="1", WORLD_SIZE="2"))
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_clock_sampler_keeps_only_the_timed_regions(tmp_path, monkeypatch):
    """bench.py starts nvidia-smi when the process starts (its first sample takes seconds on an 8-GPU box) and reports
    only the samples stamped inside the timed regions: a fake nvidia-smi that idles at 210 MHz, then runs at 1965 MHz
    with the power cap active, must give the latter."""
    import stat
    import time
    fake = tmp_path / "nvidia-smi"
    fake.write_text("""#!/usr/bin/env python3
import datetime, sys, time
t0 = time.time()
while True:
    busy = time.time() - t0 > 0.6
    now = datetime.datetime.now().strftime("%Y/%m/%d %H:%M:%S.%f")[:-3]
    print(f"{now}, 0, {1965 if busy else 210}, 1965, 500.0, 0x4, Not Active, Not Active, Not Active, {'Active' if busy else 'Not Active'}", flush=True)
    time.sleep(0.02)
""")
    fake.chmod(fake.stat().st_mode | stat.S_IEXEC)
    monkeypatch.setenv("PATH", str(tmp_path) + os.pathsep + os.environ["PATH"])
    sys.path.insert(0, ROOT)
    import bench
    cs = bench.ClockSampler(0).start()
    time.sleep(0.8)                      # set-up: idle clocks, must not be reported
    cs.mark_begin()
    time.sleep(0.4)                      # "timed regions"
    cs.stop()
    r = cs.result
    assert r["sm_mhz"] == 1965.0 and r["sm_max_mhz"] == 1965.0 and r["reasons"] == ["sw_power_cap"]
    assert r["samples"] >= 5 and r["window"] == "timed regions"
