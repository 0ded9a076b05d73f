#!/bin/bash
# First GPU calls of the next round: validates what was written after round 1's GPU budget was spent and A/B-measures the
# prepared opt-in variants.  Phases (one gpurun call each keeps every call well inside its time limit):
#   gpurun --timeout 1500 -- 'bash tools/ab_round2.sh A'   validation (pytest -m gpu, smoke) + diagnostics (cycle counters, ncu at C5)
#   gpurun --timeout 1500 -- 'bash tools/ab_round2.sh B'   small-d variants (C5 / C1): solo kernel, row-owner epilogue, lane = column M-step
#   gpurun --timeout 1500 -- 'bash tools/ab_round2.sh C'   C3 / C2 variants: truncating converter, E-step / M-step overlap
#   gpurun --timeout 1500 -- 'bash tools/ab_round2.sh D'   transform store pattern, C++ bench shapes, seeding on the tensor cores, C4
#   gpurun --gpus 2 --timeout 900 -- 'python -m pytest tests/test_kmeans_mg_gpu.py -m gpu -q; ./kmeans_mg_test 2'   the 2-rank tests (skipped on one GPU)
# No argument = all phases.  BUDGET_S (default 1300) stops starting new steps once that much wall-clock has passed.
# Everything lands in gpurun_out/r2/; `python tools/summarize_r2.py` turns the logs into one A/B table.  No step depends on another; a failure is logged and the script goes on.
set -u
PHASES=${1:-ABCD}
BUDGET_S=${BUDGET_S:-1300}
T0=$(date +%s)
OUT=gpurun_out/r2
mkdir -p "$OUT"
PH=A
phase() { PH=$1; }
HUNG=""
run() {
  name=$1; shift
  case "$PHASES" in *"$PH"*) ;; *) return 0 ;; esac
  if [ $(( $(date +%s) - T0 )) -gt "$BUDGET_S" ]; then echo "skipped (budget) $name" | tee -a "$OUT/index.log"; return 0; fi
  # a variant whose earlier step timed out (hung kernel) is not started again: the switches it set are its tag
  tag=$(echo "$*" | grep -o 'CUML_B200_[A-Z0-9_]*=[^ ]*' | sort | tr '\n' ' ')
  if [ -n "$tag" ] && echo "$HUNG" | grep -qF "|$tag|"; then echo "skipped (variant hung before) $name" | tee -a "$OUT/index.log"; return 0; fi
  t1=$(date +%s)
  echo "== [$PH] $name: $*" | tee -a "$OUT/index.log"; ( "$@" ) > "$OUT/$name.log" 2>&1; rc=$?
  echo "rc=$rc $name ($(( $(date +%s) - t1 )) s)" | tee -a "$OUT/index.log"
  if [ "$rc" = 124 ] || [ "$rc" = 137 ]; then HUNG="$HUNG|$tag|"; fi
}


phase A
# 1. correctness of everything new (fp64, callers, distributed world size 1, C-ABI as before)
run pytest_gpu timeout 900 python -m pytest tests -m gpu -x -q
run smoke timeout 300 python __graft_entry__.py smoke

# 2. MMA-issue floor of the pair kernel (cta_group::2): decides whether C3's E-step is at 84 % or 63 % of it
case "$PHASES" in *A*) [ -x tools/micro/mma_rate_pair ] || nvcc -gencode arch=compute_100a,code=sm_100a -O3 -Icuml_b200/csrc -o tools/micro/mma_rate_pair tools/micro/mma_rate_pair.cu ;; esac
run mma_rate_pair timeout 120 ./tools/micro/mma_rate_pair

# 2b. where the C3 E-step waits: role-level cycle counters of the pair kernel (CLK instantiation, CTA 0)
run clk_c3 timeout -k 10 240 env CUML_B200_DBG_CLK=1 python bench.py --workload C3 --steps 2 --warmup 3 --no-e2e --no-cpu

# 2c. the weakest roofline fractions of round 1 are the small-d configs (C5 fused kernel 0.30 of HBM, C1 0.33): where does
# the single-CTA kernel wait?  Role-level cycle counters (row-packed path included) + one full ncu capture of the
# fused kernel and the lane = row M-step kernel at C5
run clk_c5 timeout -k 10 240 env CUML_B200_DBG_CLK=1 python bench.py --workload C5 --steps 2 --warmup 3 --no-e2e --no-cpu
run clk_c5_nopack timeout -k 10 240 env CUML_B200_DBG_CLK=1 CUML_B200_PACK=0 python bench.py --workload C5 --steps 2 --warmup 3 --no-e2e --no-cpu
run clk_c1 timeout -k 10 200 env CUML_B200_DBG_CLK=1 python bench.py --workload C1 --steps 2 --warmup 3 --no-e2e --no-cpu
run ncu_full_c5 timeout 900 ncu --set full --import-source on --clock-control none -k regex:"fused_l2_argmin|accumulate_tma" -c 2 -o "$OUT/full_c5" python bench.py --workload C5 --steps 1 --warmup 0 --no-cpu --no-e2e
run ncu_full_c5_txt timeout 300 python tools/ncu_raw.py "$OUT/full_c5.ncu-rep"

phase C
# 3. E-step variants at C3 (fused kernel time is in roofline.kernel_ms)
run bench_c3_default timeout -k 10 240 python bench.py --workload C3 --steps 10 --no-e2e --no-cpu
run bench_c3_conv_trunc timeout -k 10 240 env CUML_B200_CONV_TRUNC=1 python bench.py --workload C3 --steps 10 --no-e2e --no-cpu
# parity of the truncating converter before its number means anything
run parity_conv_trunc timeout -k 10 240 env CUML_B200_CONV_TRUNC=1 python -m pytest tests/test_kmeans_gpu.py -m gpu -q -k "single_lloyd_step or dot_accuracy or regime2"

# 3b. E-step / M-step overlap on two streams (chunks:m_sms); parity at the sizes that take the overlapped path first
run parity_overlap timeout -k 10 240 env CUML_B200_OVERLAP=8:24 python -m pytest tests/test_kmeans_gpu.py -m gpu -q -k "large_property or full_size_c3"
for cfg in 8:24 8:32 16:24 4:24; do
  run bench_c3_overlap_${cfg/:/_} timeout -k 10 240 env CUML_B200_OVERLAP=$cfg python bench.py --workload C3 --steps 10 --no-e2e --no-cpu
done
run bench_c2_overlap_8_24 timeout -k 10 240 env CUML_B200_OVERLAP=8:24 python bench.py --workload C2 --steps 10 --no-e2e --no-cpu
run bench_c2_default timeout -k 10 240 python bench.py --workload C2 --steps 10 --no-e2e --no-cpu

phase B
# 4. single-CTA twin (k <= 128): parity first, then C1 / C5
run parity_solo_v2 timeout -k 10 240 env CUML_B200_SOLO_V2=1 python -m pytest tests/test_kmeans_gpu.py -m gpu -q -k "single_lloyd_step or dot_accuracy or regime2 or transform_matches"
run bench_c1_default timeout -k 10 200 python bench.py --workload C1 --steps 50 --no-e2e --no-cpu
run bench_c1_solo_v2 timeout -k 10 200 env CUML_B200_SOLO_V2=1 python bench.py --workload C1 --steps 50 --no-e2e --no-cpu
run bench_c5_default timeout -k 10 240 python bench.py --workload C5 --steps 10 --no-e2e --no-cpu
# C5 is row-packed (two rows per 128-byte operand row, block-diagonal centroids): the solo kernel takes the packed operands
# too (9 MMAs per 256 rows instead of 12), and the unpacked shape for comparison (half-empty K-block, nks = 2)
run bench_c5_solo_v2 timeout -k 10 240 env CUML_B200_SOLO_V2=1 python bench.py --workload C5 --steps 10 --no-e2e --no-cpu
# the row-owner epilogue on the measured 3xTF32 kernel alone (independent of the solo kernel): parity, C5, C1
run parity_rowown timeout -k 10 240 env CUML_B200_EPI_ROWOWN=1 python -m pytest tests/test_kmeans_gpu.py -m gpu -q -k "single_lloyd_step or regime2 or weighted_step or golden"
run bench_c5_rowown timeout -k 10 240 env CUML_B200_EPI_ROWOWN=1 python bench.py --workload C5 --steps 10 --no-e2e --no-cpu
run bench_c1_rowown timeout -k 10 200 env CUML_B200_EPI_ROWOWN=1 python bench.py --workload C1 --steps 50 --no-e2e --no-cpu
# ... and with the row-owner epilogue (4 warps per accumulator, 4 tiles in flight, no merge / named barriers): parity, C5, C1
run parity_solo_v2_rowown timeout -k 10 240 env CUML_B200_SOLO_V2=1 CUML_B200_EPI_ROWOWN=1 python -m pytest tests/test_kmeans_gpu.py -m gpu -q -k "single_lloyd_step or dot_accuracy or regime2 or weighted_step"
run bench_c5_solo_v2_rowown timeout -k 10 240 env CUML_B200_SOLO_V2=1 CUML_B200_EPI_ROWOWN=1 python bench.py --workload C5 --steps 10 --no-e2e --no-cpu
run bench_c1_solo_v2_rowown timeout -k 10 200 env CUML_B200_SOLO_V2=1 CUML_B200_EPI_ROWOWN=1 python bench.py --workload C1 --steps 50 --no-e2e --no-cpu
run bench_c5_solo_v2_nopack timeout -k 10 240 env CUML_B200_SOLO_V2=1 CUML_B200_PACK=0 python bench.py --workload C5 --steps 10 --no-e2e --no-cpu
run bench_c5_nopack timeout -k 10 240 env CUML_B200_PACK=0 python bench.py --workload C5 --steps 10 --no-e2e --no-cpu

# 4a. M-step for short rows: lane = column kernel with warp-private tables (plain CUDA, no TMA ring) vs the lane = row kernel
run parity_upd_lanecol timeout -k 10 240 env CUML_B200_UPD_LANECOL=1 python -m pytest tests/test_kmeans_gpu.py -m gpu -q -k "single_lloyd_step or weighted_step or skewed or regime2"
run bench_c5_upd_lanecol timeout -k 10 240 env CUML_B200_UPD_LANECOL=1 CUML_B200_UPD_PLAN=1 python bench.py --workload C5 --steps 10 --no-e2e --no-cpu
run bench_c5_lanecol_solo_v2 timeout -k 10 240 env CUML_B200_UPD_LANECOL=1 CUML_B200_SOLO_V2=1 python bench.py --workload C5 --steps 10 --no-e2e --no-cpu

phase D
# 4b. transform: lane-pair store pattern (DIST = 2 instantiations) -- parity, then the C4-shape probe both ways
run parity_dist_pairst timeout -k 10 240 env CUML_B200_DIST_PAIRST=1 python -m pytest tests/test_kmeans_gpu.py -m gpu -q -k "transform"
run c4_probe_default timeout -k 10 240 python tools/c4_probe.py
run c4_probe_pairst timeout -k 10 240 env CUML_B200_DIST_PAIRST=1 python tools/c4_probe.py

# 4c. the reference's C++ benchmark shapes through the C++ surface (examples/kmeans_bench.cpp)
case "$PHASES" in *D*) g++ -O2 -std=c++17 -Iinclude -I/usr/local/cuda/include examples/kmeans_bench.cpp -Lcuml_b200/lib -lcuml_b200 -L/usr/local/cuda/lib64 -lcudart -Wl,-rpath,$PWD/cuml_b200/lib -o /tmp/kmeans_bench ;; esac
run cpp_bench timeout 900 /tmp/kmeans_bench
# 4c'. the reference's multi-GPU gtest inputs through the C++ surface on a handle with an injected (one-rank) NCCL communicator
case "$PHASES" in *D*) g++ -O2 -std=c++17 -Iinclude -I/usr/local/cuda/include examples/kmeans_mg_test.cpp -Lcuml_b200/lib -lcuml_b200 -L/usr/local/cuda/lib64 -lcudart -Wl,-rpath,$PWD/cuml_b200/lib -o /tmp/kmeans_mg_test ;; esac
run cpp_mg_test timeout 600 /tmp/kmeans_mg_test

# 4d. k-means|| seeding with the min-distance updates on the tensor-core kernel: parity, then the C5 init time both ways
run parity_seed_tc timeout -k 10 240 env CUML_B200_SEED_TC=1 python -m pytest tests/test_kmeans_gpu.py tests/test_z_callers.py -m gpu -q -k "seeded or sampling"
run bench_c5_seed_tc timeout -k 10 240 env CUML_B200_SEED_TC=1 python bench.py --workload C5 --steps 5 --no-e2e --no-cpu
# (bench_c5_default above carries the CUDA-core init time in its "init" object)

# 5. the inference config (new bench workload)
run bench_c4 timeout -k 10 240 python bench.py --workload C4 --steps 3 --no-cpu
grep -h '^{' "$OUT"/bench_*.log > "$OUT/bench_lines.jsonl" 2>/dev/null
tail -n 3 "$OUT"/*.log | tail -n 120
