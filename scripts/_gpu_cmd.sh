timeout 1500 python -m pytest tests -m gpu -q --timeout 300 -x -p no:cacheprovider 2>&1 | tail -5
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; tail -c 1500 gpurun_out/bench_n1.json; tail -3 gpurun_out/bench_n1.err
