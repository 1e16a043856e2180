for ko in 15 7; do
HV_TC_KO=$ko HV_ATTN_TCGEN05=1 HV_TC_TRACE_DUMP=gpurun_out/tc_trace_ko$ko.txt timeout 120 python tools/profile_attn.py --batch 128 --iters 2 --shift 0
done
