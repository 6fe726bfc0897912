#!/bin/bash
# round 2, step 24: pass 2 with 14 digit products per K step: parity (tensor-core tests, executor, full size), then the decomposition
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_tcgauss.py tests/test_gpu_executor.py tests/test_gpu_fullsize.py -m gpu -x -q > gpurun_out/r02i_pytest.log 2>&1; tail -5 gpurun_out/r02i_pytest.log
timeout 300 python scripts/exp_round2i.py > gpurun_out/r02i_exp.json 2> gpurun_out/r02i_exp.err; tail -3 gpurun_out/r02i_exp.err; cat gpurun_out/r02i_exp.json
