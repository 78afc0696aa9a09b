set -x
timeout 600 python -m pytest tests/test_gpu_attention.py -m gpu -q --timeout 120 -x -p no:cacheprovider -k "long_prompts" 2>&1 | tail -15
timeout 1200 python -m pytest tests -m gpu -q --timeout 300 -x -p no:cacheprovider 2>&1 | tail -5
DSC_XATTN_IMPL=mma timeout 300 python scripts/microbench.py --quick --no-ref --out gpurun_out/mb_mma.jsonl 2>&1 | grep "^{" | cut -c1-175
