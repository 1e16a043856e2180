# clock trace of CTA 0 of the tcgen05 forward (library must be built with HV_NVCC_FLAGS=-DHV_TC_TRACE)
HV_ATTN_TCGEN05=1 HV_TC_TRACE_DUMP=gpurun_out/tc_trace.txt timeout 120 python tools/profile_attn.py --batch 128 --iters 2
