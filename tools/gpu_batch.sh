#!/bin/bash
# Run a list of steps on the GPU box, each under its own timeout, logs to gpurun_out/<tag>/<name>.log.
#   bash tools/gpu_batch.sh <tag> <steps-file>
# steps file: one step per line, "name|timeout_s|command"; '#' lines are comments.  A step never stops the batch.
set -u
TAG=$1; STEPS=$2
OUT=gpurun_out/$TAG
mkdir -p "$OUT"
T0=$(date +%s)
BUDGET_S=${BUDGET_S:-1500}
while IFS='|' read -r name tmo cmd; do
  case "$name" in ''|\#*) continue ;; esac
  if [ $(( $(date +%s) - T0 )) -gt "$BUDGET_S" ]; then echo "skipped (budget) $name" | tee -a "$OUT/index.log"; continue; fi
  t1=$(date +%s)
  echo "== $name: $cmd" | tee -a "$OUT/index.log"
  timeout -k 10 "$tmo" bash -c "$cmd" > "$OUT/$name.log" 2>&1; rc=$?
  echo "rc=$rc $name ($(( $(date +%s) - t1 )) s)" | tee -a "$OUT/index.log"
done < "$STEPS"
grep -h '^{' "$OUT"/bench_*.log > "$OUT/bench_lines.jsonl" 2>/dev/null
tail -n 4 "$OUT"/*.log | tail -n 150
