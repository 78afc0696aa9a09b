timeout 900 python -m pytest tests/test_gpu_pipeline.py tests/test_gpu_region_sampler.py -m gpu -q --timeout 300 -x -p no:cacheprovider 2>&1 | tail -5
timeout 300 python -c "
import __graft_entry__ as g
g.smoke(); print('smoke ok')
" 2>&1 | tail -3
