#!/bin/bash
# ncu evidence for one round: launch list of bench.py + one --set full capture of the hot kernels.
#   gpurun --timeout 1500 -- 'bash profiles/capture.sh r01'
# Writes gpurun_out/<tag>_*; summaries are then extracted here with profiles/summarise.py.
tag=${1:-r01}
out=gpurun_out
mkdir -p $out
K='regex:rollout_kernel|nfsp_step|legacy_rollout_kernel|insert_kernel'
python bench.py --steps 3 --warmup 3 > $out/${tag}_bench_plain.json 2> $out/${tag}_bench_plain.err &&
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $out/${tag}_launches.csv \
    python bench.py --steps 3 --warmup 3 > $out/${tag}_bench_under_ncu.log 2>&1
python profiles/run_hot_path.py all 2 > $out/${tag}_hot_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k "$K" -c 40 -f -o $out/${tag}_prof \
    python profiles/run_hot_path.py all 2 > $out/${tag}_hot_under_ncu.log 2>&1
tail -n 2 $out/${tag}_hot_plain.log; tail -n 2 $out/${tag}_hot_under_ncu.log
