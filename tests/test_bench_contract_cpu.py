"""The measurement contract of bench.py, checked without a GPU: the committed round-1 bench lines (profiles/) carry
every key the driver and the judge read, with consistent values, and the reference arm (`--impl reference`, the
reference's CPU execution path on a bounded sample) prints the same line shape here."""
import glob
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
             "vs_baseline", "dtype", "data", "config", "e2e"}


def _lines():
    return sorted(p for p in glob.glob(os.path.join(ROOT, "profiles", "r*_bench_*.json")) if "reference" not in p)


@pytest.mark.parametrize("path", _lines(), ids=os.path.basename)
def test_committed_bench_lines_keep_the_contract(path):
    j = json.load(open(path))
    assert BASE_KEYS | {"gpu_launches", "clocks", "roofline"} <= set(j)
    c4 = j["config"]["workload"].startswith("C4")        # inference config: a step is one predict pass over the rows
    assert (j["metric"], j["unit"]) == (("kmeans_predict_passes_per_sec", "predict passes/s") if c4 else
                                        ("kmeans_lloyd_iters_per_sec", "Lloyd iter/s")) and j["higher_is_better"] is True
    assert j["vs_baseline"] is None                      # BASELINE.md holds no published number for this metric
    assert j["warmup"] >= 3 and j["steps"] >= 1 and j["data"] == "synthetic"
    assert abs(j["value"] * j["ms_per_step"] - 1e3) < 1e-6 * 1e3          # value = 1 / time per Lloyd iteration
    assert j["config"]["workload"][:2] in ("C1", "C2", "C3", "C4", "C5") and j["config"]["l2"] == "inputs_exceed_l2"
    assert j["gpu_launches"] > 0                         # our kernels ran inside the timed region
    e = j["e2e"]
    assert {"value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"} <= set(e) and e["unit"] == j["unit"]
    assert e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0 and 0 < e["value"] < j["value"]
    c = j["clocks"]
    # round 1's 4-GPU line carries no clock sample: nvidia-smi had not produced one before its 80 ms timed region was
    # over.  bench.py now waits for the first sample and re-samples under the same load when the region is too short
    # (test_clock_sampler_* below); every other committed line has samples from inside the timed region
    if os.path.basename(path) != "r01_bench_C3_n4.json":
        assert c["samples"] >= 2 and c["sm_mhz"] > 0 and c["sm_max_mhz"] >= c["sm_mhz"]
    assert not set(c["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    r = j["roofline"]
    assert r["bound"] in ("hbm", "tensor") and r["unit"] == ("GB/s" if r["bound"] == "hbm" else "TFLOP/s")
    assert abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    # round 1 kept the per-run facts (rows per GPU, engine, communicator) in `config`; since round 2 `config` names the
    # workload only (same at every N, so the driver's same_config check holds) and the rest lives in `run`
    rows = j["config"]["rows_per_gpu"] if "rows_per_gpu" in j["config"] else j["run"]["rows_per_gpu"]
    n, d, k = rows, j["config"]["d"], j["config"]["k"]
    # achieved = algorithmic work per launch (2nkd flop, 4nd + 4n bytes) / measured launch duration
    assert r["algorithmic_flops_per_launch"] == 2.0 * n * k * d and r["algorithmic_bytes_per_launch"] == 4.0 * n * d + 4.0 * n
    work = r["algorithmic_flops_per_launch"] / 1e12 if r["bound"] == "tensor" else r["algorithmic_bytes_per_launch"] / 1e9
    assert abs(r["achieved"] - work / (r["kernel_ms"] * 1e-3)) <= 1e-6 * r["achieved"]
    if j["n_gpus"] == 1:
        b = j["cpu_baseline"]
        assert b["kind"] == "reference" and b["cores"] >= 1 and b["value"] > 0 and b["unit"] == j["unit"] and b["sample"]
        # ncu DRAM traffic of the dominant kernel, per launch like `achieved`: within 2 % of the algorithmic bytes
        # (C4 re-streams X once per centroid tile from L2: its DRAM traffic is a multiple of the algorithmic bytes)
        if r.get("traffic") and not c4:
            assert abs(r["traffic"] / r["algorithmic_bytes_per_launch"] - 1.0) < 0.02


def test_scaling_series_is_whole_job_throughput():
    by_n = {}
    for p in _lines():
        j = json.load(open(p))
        if j["config"]["workload"].startswith("C3"):
            by_n[j["n_gpus"]] = j
    assert sorted(by_n) == [1, 2, 4, 8]
    for g, j in by_n.items():
        assert j["scaling"] == "strong" and j["config"]["n"] == 100_000_000     # total work fixed as N grows
        rows = j["config"]["rows_per_gpu"] if "rows_per_gpu" in j["config"] else j["run"]["rows_per_gpu"]
        assert rows * g >= j["config"]["n"]
    assert by_n[8]["value"] / by_n[1]["value"] >= 0.85 * 8                       # the north-star scaling bar


def test_reference_arm_prints_the_contract_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "C1",
                        "--n", "100000", "--steps", "1", "--warmup", "3"], capture_output=True, text=True, timeout=600,
                       cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1                                # ONE JSON line
    j = json.loads(lines[0])
    assert BASE_KEYS | {"impl", "cpu_baseline"} <= set(j) and j["impl"] == "reference"
    assert j["metric"] == "kmeans_lloyd_iters_per_sec" and j["value"] > 0 and j["higher_is_better"] is True
    assert j["e2e"] == {"value": j["value"], "unit": j["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert j["cpu_baseline"]["kind"] == "reference" and j["cpu_baseline"]["value"] == j["value"]
    assert j["cpu_baseline"]["cores"] >= 1 and "sklearn" in j["cpu_baseline"]["sample"]


def test_reference_arm_other_ranks_exit_without_work():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"],
                       capture_output=True, text=True, timeout=120, cwd=ROOT, env=env)
    assert r.returncode == 0 and not [ln for ln in r.stdout.splitlines() if ln.startswith("{")]


def _fake_nvidia_smi(tmp_path, delay_s):
    """an `nvidia-smi` that starts slowly (as on a multi-GPU box) and then prints one CSV sample every 20 ms"""
    exe = tmp_path / "nvidia-smi"
    exe.write_text("#!/bin/sh\nsleep %s\nwhile true; do echo \"0, 1695, 1965, 612.3, Not Active, Not Active, Not Active, Active\"; "
                   "sleep 0.02; done\n" % delay_s)
    exe.chmod(0o755)
    return str(tmp_path)


def test_clock_sampler_waits_for_the_first_sample_and_windows(tmp_path, monkeypatch):
    import time
    sys.path.insert(0, ROOT)
    import bench
    monkeypatch.setenv("PATH", _fake_nvidia_smi(tmp_path, 0.3) + os.pathsep + os.environ["PATH"])
    s = bench.ClockSampler(0)
    s.start()
    t0 = time.time()
    s.wait_first()
    assert s.lines and 0.2 < time.time() - t0 < 3.0        # blocked until the slow start was over
    t_begin = time.time()
    time.sleep(0.25)
    t_end = time.time()
    assert s.inside(t_begin, t_end) >= 2
    assert s.inside(t_begin - 100.0, t_begin - 99.0) == 0    # a window nothing fell into (the probe trigger)
    c = s.stop(t_begin, t_end)
    assert c["samples"] >= 2 and c["sm_mhz"] == 1695.0 and c["sm_max_mhz"] == 1965.0 and c["reasons"] == ["sw_power_cap"]


def test_clock_sampler_without_nvidia_smi(tmp_path, monkeypatch):
    sys.path.insert(0, ROOT)
    import bench
    monkeypatch.setenv("PATH", str(tmp_path))                # no nvidia-smi anywhere
    s = bench.ClockSampler(0)
    s.start()
    s.wait_first(timeout=0.2)
    assert s.stop(0.0, 1.0) == dict(sm_mhz=None, sm_max_mhz=None, reasons=[], samples=0)
