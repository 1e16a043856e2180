"""Print one line per attention row of a tools/bench_kernels.py JSON."""
import json
import sys

for fn in sys.argv[1:]:
    print(fn)
    for r in json.load(open(fn)):
        if r["kernel"] != "window_attn":
            continue
        print("  C%-4d shift %d %-8s tau=%-5s fwd %.3f ms (%.2f)  bwd %.3f ms (%.2f)  fwd+bwd %.2f  %.1f Mwin/s" % (
            r["C"], r["shift"], r["dtype"], r.get("tau", "?"), r["fwd_ms"], r["frac_fwd"], r["bwd_ms"], r["frac_bwd"],
            r["frac_fwdbwd"], r["windows_per_s"] / 1e6))
