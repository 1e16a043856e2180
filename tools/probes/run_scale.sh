for b in 64 128 256 512; do for v in 0 1; do
  HV_ATTN_TCGEN05_BWD=$v timeout 300 python tools/bench_kernels.py --batch $b --only attn0 --iters 20 --json gpurun_out/kb_sc.json > /dev/null 2>&1
  echo "batch $b bwd_variant $v: $(python tools/kb_summary.py gpurun_out/kb_sc.json | sed -n 2,2p)"
done; done
