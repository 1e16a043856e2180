python -m pytest tests -m gpu -x -q -k "attention or attn" 2>&1 | tail -2
timeout 300 python tools/bench_kernels.py --batch 256 --only attn --iters 30 --json gpurun_out/kb_b256_f2.json > /dev/null 2>&1
python tools/kb_summary.py gpurun_out/kb_b256_f2.json
