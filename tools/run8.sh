timeout 600 python -m pytest tests/test_kmeans_gpu.py -x -q -m gpu 2>&1 | tail -3
run() { env "$@" timeout 300 python bench.py --workload $W --steps 10 --warmup 3 --no-cpu --no-e2e 2>&1 | tail -1 > gpurun_out/tmp.json; python - <<PY
import json
d=json.load(open("gpurun_out/tmp.json")); r=d["roofline"]
print("$W $*", "ms/step", round(d["ms_per_step"],3), "fused", round(r["kernel_ms"],3), "update_ms", round(r["update_kernel_ms"],3), "GBs", round(r["update_kernel_hbm_gbs"]), d["clocks"])
PY
}
W=C3
run CUML_B200_UPDATE_OWNER=0
run A=1
run CUML_B200_UPDATE_OWNER=0
run A=1
W=C5
run A=1
W=C1
run A=1
