python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py > gpurun_out/bench_s2d_n1.json 2> gpurun_out/bench_s2d_n1.err; echo "bench rc=$?"; cut -c1-400 gpurun_out/bench_s2d_n1.json
