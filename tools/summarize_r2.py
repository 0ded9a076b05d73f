#!/usr/bin/env python
"""One table from the logs of tools/ab_round2.sh (gpurun_out/r2/):

    python tools/summarize_r2.py [gpurun_out/r2]

* every `bench_*.log` that holds a bench.py JSON line: workload, Lloyd iter/s, ms per step, fused-kernel ms, M-step ms,
  roofline fraction, SM clock and throttle reasons, the k-means|| init seconds when present;
* every `parity_*.log` / `pytest_gpu.log`: the pytest summary line;
* `index.log` return codes of the remaining steps (cycle counters, ncu, micro-benchmarks, C++ programs).
Variants are grouped by workload so that each A/B reads top to bottom against its `*_default` line.
"""
import glob
import json
import os
import re
import sys


def bench_line(path):
    for line in open(path, errors="replace"):
        if line.startswith("{") and '"metric"' in line:
            try:
                return json.loads(line)
            except json.JSONDecodeError:
                pass
    return None


def main():
    out = sys.argv[1] if len(sys.argv) > 1 else os.path.join("gpurun_out", "r2")
    rows = []
    for path in sorted(glob.glob(os.path.join(out, "bench_*.log"))):
        name = os.path.basename(path)[len("bench_"):-len(".log")]
        j = bench_line(path)
        if j is None:
            rows.append((name.split("_")[0], name, None))
            continue
        rows.append((j["config"]["workload"].split(":")[0], name, j))
    rows.sort(key=lambda r: (r[0], "default" not in r[1], r[1]))
    print(f"{'workload':8} {'variant':28} {'value':>10} {'ms/step':>9} {'fused ms':>9} {'M-step ms':>9} {'roof':>6} {'MHz':>5}  notes")
    for wl, name, j in rows:
        if j is None:
            print(f"{wl:8} {name:28} {'(no bench line: see the log)':>10}")
            continue
        r = j.get("roofline") or {}
        c = j.get("clocks") or {}
        notes = ",".join(c.get("reasons") or [])
        if "init" in j:
            notes += f" init {j['init']['seconds']:.3f}s"
        frac = r.get("frac")
        print(f"{wl:8} {name:28} {j['value']:10.2f} {j['ms_per_step']:9.3f} {r.get('kernel_ms', float('nan')):9.3f} "
              f"{r.get('update_kernel_ms', float('nan')):9.3f} {frac if frac is None else round(frac, 3)!s:>6} "
              f"{c.get('sm_mhz', 0):5.0f}  {notes}")
    print()
    for path in sorted(glob.glob(os.path.join(out, "parity_*.log")) + glob.glob(os.path.join(out, "pytest_gpu.log"))):
        text = open(path, errors="replace").read()
        m = re.findall(r"^(?:=+ )?(\d+ (?:passed|failed|error)[^\n]*)", text, flags=re.M)
        print(f"{os.path.basename(path)[:-4]:28} {m[-1].strip('= ') if m else '(no pytest summary)'}")
    print()
    idx = os.path.join(out, "index.log")
    if os.path.exists(idx):
        for line in open(idx):
            if line.startswith("rc=") and not re.match(r"rc=\d+ (bench_|parity_|pytest_gpu)", line):
                print(line.rstrip())
            elif line.startswith("skipped"):
                print(line.rstrip())


if __name__ == "__main__":
    main()
