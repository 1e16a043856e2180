python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_bwd.log 2>&1; echo "pytest exit=$?" >> gpurun_out/pytest_gpu_bwd.log; tail -5 gpurun_out/pytest_gpu_bwd.log
for v in 1; do
HV_ATTN_TCGEN05_BWD=$v timeout 300 python tools/bench_kernels.py --batch 256 --only attn --iters 30 --json gpurun_out/kb_b256_bwd$v.json > /dev/null 2>&1
python tools/kb_summary.py gpurun_out/kb_b256_bwd$v.json
done
