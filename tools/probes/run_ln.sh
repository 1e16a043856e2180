python -m pytest tests -m gpu -x -q -k "ln_residual or block or patch_merging or model_tiny" 2>&1 | tail -2
timeout 300 python tools/bench_kernels.py --batch 256 --only ln --iters 30 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    try: r = json.loads(l)
    except Exception: continue
    print('  rows %8d C %4d y %s res %s fwd %.3f ms (%.2f)  bwd %.3f ms (%.2f)' % (r['rows'], r['C'], r['y'], r['res'], r['fwd_ms'], r['frac_fwd'], r['bwd_ms'], r['frac_bwd']))
"
