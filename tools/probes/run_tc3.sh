HV_ATTN_TCGEN05=1 HV_TC_TRACE_DUMP=gpurun_out/tc_trace.txt timeout 120 python tools/profile_attn.py --batch 128 --iters 2 --shift 0
HV_ATTN_TCGEN05=1 timeout 300 python tools/bench_kernels.py --batch 128 --only attn0 --json gpurun_out/kb_tcv1.json > /dev/null 2>&1
python tools/kb_summary.py gpurun_out/kb_tcv1.json
