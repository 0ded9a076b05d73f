timeout 300 python -m pytest tests/test_kmeans_gpu.py -x -q -m gpu 2>&1 | tail -3
run() { env "$@" CUML_B200_UPD_PLAN=1 timeout 300 python bench.py --workload $W --steps 6 --warmup 3 --no-cpu --no-e2e 2>&1 | tail -2 > gpurun_out/tmp.txt; head -1 gpurun_out/tmp.txt | cut -c1-120; tail -1 gpurun_out/tmp.txt > gpurun_out/tmp.json; python - <<PY
import json
d=json.load(open("gpurun_out/tmp.json")); r=d["roofline"]
print("$W $*", "ms/step", round(d["ms_per_step"],3), "fused", round(r["kernel_ms"],3), "update_ms", round(r["update_kernel_ms"],3), "GBs", round(r["update_kernel_hbm_gbs"]))
PY
}
W=C3
run A=1
run CUML_B200_OWNER_TR=256 CUML_B200_OWNER_NL=4
run CUML_B200_OWNER_TR=192
run CUML_B200_OWNER_TR=160
run CUML_B200_OWNER_TR=128
run CUML_B200_OWNER_TR=224
W=C2
run A=1
run CUML_B200_OWNER_TR=128
run CUML_B200_OWNER_TR=256
run CUML_B200_OWNER_VEC=2
