python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --no-cpu > gpurun_out/bench_s2e_n1.json 2> gpurun_out/bench_s2e_n1.err; echo "bench rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/bench_s2e_n1.json')); print({k:d[k] for k in ('value','ms_per_step','e2e')})"
