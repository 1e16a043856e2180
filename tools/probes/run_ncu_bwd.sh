HV_ATTN_TCGEN05_BWD=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:wattn_tc64_bwd_kernel -s 1 -c 1 -o gpurun_out/tc64bwd_a -f python tools/profile_attn.py --batch 128 --iters 2 > gpurun_out/ncu_tc64bwd.log 2>&1
tail -2 gpurun_out/ncu_tc64bwd.log
