# ncu --set full of the kernels matching $1 (demangled regex), launch skip $2, count $3 -> gpurun_out/$4.ncu-rep
set -x
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"$1" -s $2 -c $3 -o gpurun_out/$4 -f python scripts/prof_chunk.py 8 2 > gpurun_out/$4.log 2>&1; tail -2 gpurun_out/$4.log
