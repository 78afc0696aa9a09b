set -x
timeout 1200 python -m pytest tests -m gpu -q --timeout 300 -x -p no:cacheprovider 2>&1 | tail -5
timeout 600 python scripts/microbench.py --quick --no-ref --out gpurun_out/mb_quick.jsonl 2>&1 | grep "^{" | cut -c1-175
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; tail -c 2500 gpurun_out/bench_n1.json
