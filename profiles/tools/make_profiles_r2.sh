#!/bin/bash
# Runs on the GPU box: plain bench, ncu launch list of the same command, ncu --set full of one step's kernels.
set -x
python bench.py --steps 2 --warmup 3 --no-graph --no-cpu-baseline > gpurun_out/prof_plain_bench.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 400 --csv --log-file gpurun_out/prof_launches.csv python bench.py --steps 2 --warmup 3 --no-graph --no-cpu-baseline > gpurun_out/prof_ncu_bench.log 2>&1
python profiles/tools/prof_kernels.py > gpurun_out/prof_plain_step.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"drosfm" --kernel-id :::1 -o gpurun_out/prof_full -f python profiles/tools/prof_kernels.py > gpurun_out/prof_ncu_step.log 2>&1
tail -2 gpurun_out/prof_ncu_step.log
