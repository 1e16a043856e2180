for v in 0 1; do
HV_ATTN_TCGEN05=$v timeout 300 python tools/bench_kernels.py --batch 256 --only attn --iters 30 --json gpurun_out/kb_b256_v$v.json > /dev/null 2>&1
python tools/kb_summary.py gpurun_out/kb_b256_v$v.json | head -8
done
