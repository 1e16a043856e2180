for hg in 3 2 1; do echo "HG=$hg"; HV_ATTN_HEADS_PER_CTA=$hg timeout 200 python tools/bench_kernels.py --batch 128 --only attn --json gpurun_out/kb_hg$hg.json > /dev/null 2>&1; python - <<PY
import json
for r in json.load(open("gpurun_out/kb_hg$hg.json"))[:6]:
    print(r["C"], r["shift"], "fwd %.3f ms (%.2f)  bwd %.3f ms (%.2f)" % (r["fwd_ms"], r["frac_fwd"], r["bwd_ms"], r["frac_bwd"]))
PY
done
