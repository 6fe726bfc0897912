#!/bin/bash
# round 2, steps 24+: tensor-core Gaussian experiments: parity (tensor-core tests), then the decomposition script
set -x
mkdir -p gpurun_out
T=${TAG:-r02i}
timeout 600 python -m pytest tests/test_gpu_tcgauss.py ${EXTRA_TESTS} -m gpu -x -q > gpurun_out/${T}_pytest.log 2>&1; tail -5 gpurun_out/${T}_pytest.log
timeout 300 python scripts/exp_round2i.py > gpurun_out/${T}_exp.json 2> gpurun_out/${T}_exp.err; tail -3 gpurun_out/${T}_exp.err; cat gpurun_out/${T}_exp.json
