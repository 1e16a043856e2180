bash tools/probes/run_trace.sh
HV_ATTN_TCGEN05=1 timeout 300 python tools/bench_kernels.py --batch 128 --only attn --json gpurun_out/kb_tcv1.json > /dev/null 2>&1
python tools/kb_summary.py gpurun_out/kb_tcv1.json
HV_ATTN_TCGEN05=1 python -m pytest tests -m gpu -x -q -k "attn or attention or block" 2>&1 | tail -2
