#!/bin/bash
# Clock trace / knock-out runs of the N = 256 backward kernel (wattn_tc256_bwd.cu built with -DHV_TC256_TRACE -DHV_TC256_KO):
#   tools/build_variant.sh tr "-DHV_TC256_TRACE -DHV_TC256_KO" wattn_tc256_bwd.cu
#   KOS="0 13" bash tools/probes/tc256_trace.sh        # on the GPU box; dumps gpurun_out/tc256_trace_ko<bits>.txt
# Knock-out bits: 1 no output MMAs | 2 no d(bias) MMAs | 4 no staging stores | 8 no softmax math | 16 no G' store.
# Columns of a dump row (one row per item of CTA 0, clock64 relative to the first S / dP issue): A_sfree A_issued s_sdp s_ld
# s_math s_stfree s_staged B_staged B_accfree B_issued kv_acc kv_free kv_written q_acc st_written st_released.
for ko in ${KOS:-0 13}; do
  HV_TC256_KO=$ko HV_TC256_TRACE_DUMP=gpurun_out/tc256_trace_ko$ko.txt HV_SWIN_LIB=$PWD/hierarchical_vision_b200/libhv_swin_tr.so \
    timeout 100 python tools/profile_attn.py --batch 128 --res 64 --C 128 --heads 4 --ws 16 --shift 0 --iters 2
done
