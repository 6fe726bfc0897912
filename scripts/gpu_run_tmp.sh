set -x
mkdir -p gpurun_out
timeout 500 python scripts/exp_round2h.py > gpurun_out/r02k_exp.json 2> gpurun_out/r02k_exp.err; echo rc=$?; tail -5 gpurun_out/r02k_exp.err
python -c "
import json; d=json.load(open('gpurun_out/r02k_exp.json'))
for k,v in d.items(): print(k, v)"
python -m pytest tests/test_gpu_masks.py tests/test_gpu_executor.py tests/test_gpu_configs.py -m gpu -x -q > gpurun_out/r02k_pytest.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/r02k_pytest.log
