timeout 600 python -m pytest tests/test_gpu_attention.py -m gpu -q --timeout 120 -x -p no:cacheprovider -k "ip_adapter" 2>&1 | tail -8
