HV_ATTN_TCGEN05_BWD=1 HV_TC_BTRACE_DUMP=gpurun_out/tc_btrace.txt timeout 120 python tools/profile_attn.py --batch 128 --iters 2 --shift ${SHIFT:-0}
