# forward variants side by side: parity tests first, then the stage-shape microbenchmark with each forward kernel
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_tc.log 2>&1; echo "pytest exit=$?" >> gpurun_out/pytest_gpu_tc.log; tail -3 gpurun_out/pytest_gpu_tc.log
for v in 0 1; do
  HV_ATTN_TCGEN05=$v timeout 300 python tools/bench_kernels.py --batch 128 --only attn --json gpurun_out/kb_tcv$v.json > /dev/null 2>&1
  python tools/kb_summary.py gpurun_out/kb_tcv$v.json | head -9
done
