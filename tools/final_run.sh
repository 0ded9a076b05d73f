#!/bin/bash
# Round-end measurement pass on one B200: GPU tests, smoke, bench lines, ncu launch list + full captures.
# Outputs under gpurun_out/final/ (copied into profiles/ by hand).
set -u
O=gpurun_out/final; mkdir -p $O
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -4 > $O/pytest_gpu.txt; cat $O/pytest_gpu.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee $O/smoke.txt
timeout 900 python bench.py --workload C3 2>&1 | tail -1 > $O/bench_C3_n1.json
for w in C1 C2 C5; do timeout 600 python bench.py --workload $w 2>&1 | tail -1 > $O/bench_${w}_n1.json; done
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 2>&1 | tail -1 > $O/bench_C3_reference_arm.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_c3.csv python bench.py --workload C3 --steps 2 --warmup 3 --no-cpu --no-e2e > $O/launches_c3.log 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:"fused_l2_argmin|accumulate_owner" -c 2 -o $O/full_c3 python bench.py --workload C3 --steps 1 --warmup 0 --no-cpu --no-e2e > $O/full_c3.log 2>&1
timeout 600 ncu --set full --clock-control none -k regex:"fused_l2_argmin|accumulate_owner" -c 2 -o $O/full_c2 python bench.py --workload C2 --steps 1 --warmup 0 --no-cpu --no-e2e > $O/full_c2.log 2>&1
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/final/bench_*.json")):
    try:
        d=json.load(open(f)); r=d.get("roofline",{})
        print(f.split("/")[-1], "value", round(d["value"],3), "ms", round(d["ms_per_step"],3), "e2e", d.get("e2e",{}) and round(d["e2e"]["value"],2), "fused", r.get("kernel_ms"), "upd", r.get("update_kernel_ms"), "frac", r.get("frac"), "cpu", (d.get("cpu_baseline") or {}).get("value"), d.get("clocks"))
    except Exception as e:
        print(f, "ERR", e, open(f).read()[-300:])
PY
