for sw in "1 1" "3 3" "5 3"; do set -- $sw
  timeout 600 python bench.py --steps $1 --warmup $2 --no-cpu-baseline 2>/dev/null | python -c "
import sys, json
r = json.loads(sys.stdin.read().strip().splitlines()[-1]); f = r['roofline']
print('steps', r['steps'], 'warmup', r['warmup'], 'value %.2f e2e %.2f' % (r['value'], r['e2e']['value']), 'call %.1f (median %.1f) stats %.1f fwd %.1f us frac %.3f all16 %.3f ms' % (f['avg_ms_call']*1e3, f['median_ms_call']*1e3, f['avg_ms_stats_alone']*1e3, f['avg_ms_forward_alone']*1e3, f['frac'], f['all_16_layers']['ms_per_unet_step']), r['clocks'])
"
done
DSC_BENCH_NO_GRAPH=1 timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 20000 -c 2600 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
tail -2 gpurun_out/ncu_bench.log | cut -c1-300; wc -l gpurun_out/launches.csv
