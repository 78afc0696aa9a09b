set -x
T=/root/repo/diffusionspatialcontrol_b200/libdsc_b200_trace.so
DSC_LIB=$T DSC_XATTN_IMPL=tc5 timeout 120 python scripts/tc5_trace.py 16 4096 40 stats > gpurun_out/trace_x4_stats.txt 2>&1
DSC_LIB=$T DSC_XATTN_IMPL=tc5 timeout 120 python scripts/tc5_trace.py 16 4096 40 > gpurun_out/trace_x4_fwd.txt 2>&1
DSC_LIB=$T DSC_XATTN_IMPL=tc5 timeout 120 python scripts/tc5_trace.py 16 1024 80 > gpurun_out/trace_x4_fwd80.txt 2>&1
for combo in "auto auto" "tc5 tc5" "mma mma" "tc5 mma"; do set -- $combo
  echo "== fwd=$1 stats=$2"
  DSC_XATTN_IMPL=$1 DSC_XATTN_STATS_IMPL=$2 timeout 300 python scripts/microbench.py --quick --no-ref --shapes 1024x80,4096x40 --out gpurun_out/mb_$1_$2.jsonl 2>&1 | grep "^{" | cut -c40-175
done
timeout 900 python -m pytest tests/test_gpu_attention.py -m gpu -q --timeout 300 -x -p no:cacheprovider 2>&1 | tail -5
