#!/bin/bash
# ncu --set full of ONE launch of the fused pass 2 (and one of pass 1), with source counters
set -x
mkdir -p gpurun_out
timeout 300 python scripts/prof_fused.py > gpurun_out/r02j_prof_plain.log 2>&1; tail -1 gpurun_out/r02j_prof_plain.log
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:'tcg_axis1|tcg_axis0' -c 2 -o gpurun_out/r02j_fused -f python scripts/prof_fused.py > gpurun_out/r02j_ncu.log 2>&1; tail -2 gpurun_out/r02j_ncu.log
ls -la gpurun_out
