set -x
timeout 300 python scripts/prof_chunk.py 16 3 > gpurun_out/r02c_prof_plain.log 2>&1; tail -1 gpurun_out/r02c_prof_plain.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled -k regex:amt:: -c 2000 --csv --log-file gpurun_out/r02c_launches.csv python scripts/prof_chunk.py 16 2 > gpurun_out/r02c_ncu_launches.log 2>&1; tail -1 gpurun_out/r02c_ncu_launches.log
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:'amt::(ccl_|region_|sel_|rank_|map_kernel|relabel|exact_eval|dx_patch|present|scan_kernel|plan_|otsu|acc_init)' -s 72 -c 80 -o gpurun_out/r02c_rest -f python scripts/prof_chunk.py 16 2 > gpurun_out/r02c_ncu_rest.log 2>&1; tail -2 gpurun_out/r02c_ncu_rest.log
python scripts/stage_profile.py 32 > gpurun_out/r02c_stage_profile.json 2>/dev/null; head -c 700 gpurun_out/r02c_stage_profile.json
