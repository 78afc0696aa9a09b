WD=/root/repo/diffusionspatialcontrol_b200/libdsc_b200_wd.so
for c in "fwd 16 1024 80" "fwd 2 1024 80" "fwd 3 200 80 40" "fwd 8 2304 80"; do
  DSC_XATTN_IMPL=tc5 DSC_LIB=$WD timeout 60 python scripts/tc5_debug.py $c 2>&1 | grep -E "rel-L2|bad rows|Error|error|watchdog" | tail -2
done
timeout 300 python scripts/microbench.py --quick --no-ref --shapes 1024x80,2304x80 --out gpurun_out/mb_d80.jsonl 2>&1 | grep "^{" | cut -c1-175
timeout 600 python -m pytest tests/test_gpu_attention.py -m gpu -q --timeout 120 -x -p no:cacheprovider 2>&1 | tail -3
