# quick loop: parity tests, stage profile, device-resident bench
set -x
python -m pytest tests/test_gpu_tcgauss.py tests/test_gpu_executor.py tests/test_gpu_fullsize.py -x -q -m gpu 2>&1 | tail -3
for c in ${CHUNKS:-8 16}; do python bench.py --steps 3 --warmup 3 --fovs 128 --no-e2e --no-cpu --no-modes --chunk $c 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('chunk',d['config'].get('chunk_fovs'),'value',d['value'],'ms/step',d['ms_per_step'])"; done
