bash tools/probes/run_btrace.sh
for ko in 0 8 2 1 16 3 31; do
  echo "KO=$ko"; HV_TC_KO=$ko HV_ATTN_TCGEN05_BWD=1 timeout 300 python tools/bench_kernels.py --batch 128 --only attn0 --iters 20 --json gpurun_out/kb_ko.json > /dev/null 2>&1
  python tools/kb_summary.py gpurun_out/kb_ko.json | sed -n 2,2p
done
