HV_ATTN_TCGEN05_BWD=1 timeout 300 python tools/bench_kernels.py --batch 256 --only attn0 --iters 30 --json gpurun_out/kb_b256_bwd1.json > /dev/null 2>&1
python tools/kb_summary.py gpurun_out/kb_b256_bwd1.json
