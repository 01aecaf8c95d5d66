#!/bin/bash
# Regenerates every profiles/r02_*.csv of this round from ncu (run on a B200 box: `gpurun -- bash tools/profile_all.sh`).
# Each ncu pass runs only after its program has exited 0 without ncu.  Raw reports go to gpurun_out/ (scratch); the
# summaries that DESIGN.md and bench.py cite are written by tools/profile_summary.py into profiles/ (tracked).
set -u
OUT=gpurun_out
mkdir -p $OUT profiles
PY=python

# 1. per-kernel full captures at the bench shapes (one launch of each, second round)
$PY tools/prof_shapes.py > $OUT/prof_shapes_plain.log 2>&1 || { echo "prof_shapes failed"; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:"k1v2|ln_lora_u|ln_fwd|k2_" -s 13 -c 13 \
    -o $OUT/r02_shapes -f $PY tools/prof_shapes.py > $OUT/r02_shapes_ncu.log 2>&1
ncu -i $OUT/r02_shapes.ncu-rep --page raw --csv > $OUT/r02_shapes_raw.csv 2>/dev/null

# 2. launch list of ONE routed step of the bench (NVTX range "timed")
$PY bench.py --steps 1 --warmup 3 --skip-cpu-baseline --extras none > $OUT/r02_bench_for_launch_list.json 2>/dev/null \
    || { echo "bench failed"; exit 1; }
ncu --nvtx --nvtx-include "timed/" --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file $OUT/r02_launch_list_one_step.csv $PY bench.py --steps 1 --warmup 3 --skip-cpu-baseline --extras none \
    > $OUT/r02_launch_list_ncu.log 2>&1

# 2b. K3 alone at the training shapes: whole-call event time + per-launch durations
$PY tools/prof_k3.py > $OUT/r02_k3_plain.log 2>&1 || { echo "prof_k3 failed"; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none --csv -k regex:"k1v2|k1_qv|k3_" --log-file $OUT/r02_k3_launches.csv \
    $PY tools/prof_k3.py > $OUT/r02_k3_ncu.log 2>&1

# 3. the attention kernel and the encoder SDPA beside it
$PY tools/dev_check.py attn > $OUT/r02_attn_devcheck.log 2>&1

$PY tools/profile_summary.py $OUT profiles
