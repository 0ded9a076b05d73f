#!/bin/bash
# Round-end measurement pass on one B200: GPU tests, smoke, bench lines, reference arm, ncu launch list + full captures.
# Outputs under gpurun_out/final/ (copied into profiles/ as r02_* by hand).
set -u
O=gpurun_out/final; mkdir -p $O
timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -6 > $O/pytest_gpu.txt; cat $O/pytest_gpu.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee $O/smoke.txt
timeout 900 python bench.py --steps 20 --warmup 5 2>&1 | tail -1 > $O/bench_C3_n1.json
timeout 600 python bench.py --workload C1 --steps 50 --warmup 5 2>&1 | tail -1 > $O/bench_C1_n1.json
timeout 600 python bench.py --workload C2 --steps 20 --warmup 5 2>&1 | tail -1 > $O/bench_C2_n1.json
timeout 600 python bench.py --workload C4 --steps 3 --warmup 3 2>&1 | tail -1 > $O/bench_C4_n1.json
timeout 600 python bench.py --workload C5 --steps 20 --warmup 5 2>&1 | tail -1 > $O/bench_C5_n1.json
timeout 900 python bench.py --impl reference --steps 20 --warmup 5 2>&1 | tail -1 > $O/bench_C3_reference_arm.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_c3.csv python bench.py --workload C3 --steps 2 --warmup 3 --no-cpu --no-e2e --no-parity > $O/launches_c3.log 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:"fused_l2_argmin|accumulate_owner" -s 8 -c 2 -o $O/full_c3 python bench.py --workload C3 --steps 2 --warmup 3 --no-cpu --no-e2e --no-parity > $O/full_c3.log 2>&1
timeout 600 ncu --set full --clock-control none -k regex:"fused_l2_argmin_tsp" -s 4 -c 1 -o $O/full_c5 python bench.py --workload C5 --steps 2 --warmup 3 --no-cpu --no-e2e --no-parity > $O/full_c5.log 2>&1
timeout 600 ncu --set full --clock-control none -k regex:"fused_l2_argmin" -s 3 -c 1 -o $O/full_c4 python bench.py --workload C4 --steps 1 --warmup 3 --no-cpu --no-e2e --no-parity > $O/full_c4.log 2>&1
for f in full_c3 full_c5 full_c4; do python tools/ncu_raw.py $O/$f.ncu-rep > $O/$f.txt 2>&1; done
rm -f $O/full_c5.ncu-rep $O/full_c4.ncu-rep
ls -la $O
