for ko in 0 15; do
  echo "KO=$ko"; HV_TC_KO=$ko HV_ATTN_TCGEN05=1 timeout 300 python tools/bench_kernels.py --batch 128 --only attn0 --iters 20 --json gpurun_out/kb_ko.json > /dev/null 2>&1
  python tools/kb_summary.py gpurun_out/kb_ko.json | sed -n 2,2p
done
HV_ATTN_TCGEN05=1 python -m pytest tests -m gpu -x -q -k "attn or attention or block" 2>&1 | tail -2
bash tools/probes/run_ko2.sh
