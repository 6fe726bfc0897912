set -x
SWEEP_WARM=0 SWEEP_ITERS=1 timeout 800 ncu --set full --clock-control none --import-source on -k regex:'dog_strip' -o gpurun_out/prof_strip_s2c -f python scripts/dog_sweep.py 32 > gpurun_out/ncu_s2c.log 2>&1; echo "ncu rc=$?"
tail -5 gpurun_out/ncu_s2c.log
