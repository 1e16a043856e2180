python -m pytest tests -m gpu -x -q -k "gelu" 2>&1 | tail -2
timeout 300 python tools/bench_kernels.py --batch 256 --only gelu --iters 30 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    try: r = json.loads(l)
    except Exception: continue
    print('  rows %8d cols %4d fwd %.3f ms (%.2f)  bwd %.3f ms (%.2f)' % (r['rows'], r['cols'], r['fwd_ms'], r['frac_fwd'], r['bwd_ms'], r['frac_bwd']))
"
