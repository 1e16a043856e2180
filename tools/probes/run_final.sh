set -x
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_final.log 2>&1; echo "exit=$?" >> gpurun_out/pytest_gpu_final.log; tail -3 gpurun_out/pytest_gpu_final.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/smoke_final.log 2>&1; tail -2 gpurun_out/smoke_final.log
python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; tail -c 600 gpurun_out/bench_final.json
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_final.json 2> gpurun_out/bench_ref_final.err; tail -c 400 gpurun_out/bench_ref_final.json
python tools/step_profile.py --batch 256 --steps 3 > gpurun_out/step_profile_final.txt 2>&1; head -12 gpurun_out/step_profile_final.txt
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches_final.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/ncu_launches_final.log 2>&1; tail -2 gpurun_out/ncu_launches_final.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:wattn -s 2 -c 2 -o gpurun_out/attn_final -f python tools/profile_attn.py --batch 256 --iters 2 > gpurun_out/ncu_attn_final.log 2>&1; tail -1 gpurun_out/ncu_attn_final.log
HV_ATTN_TCGEN05_BWD=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:wattn_tc64_bwd_kernel -s 1 -c 1 -o gpurun_out/attn_tcbwd_final -f python tools/profile_attn.py --batch 256 --iters 2 > gpurun_out/ncu_attn_tcbwd_final.log 2>&1; tail -1 gpurun_out/ncu_attn_tcbwd_final.log
