set -x
python scripts/dog_sweep.py 32 > gpurun_out/dog_sweep_a.log 2>&1; echo "sweep rc=$?"
tail -12 gpurun_out/dog_sweep_a.log
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_s2b.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu_s2b.log
