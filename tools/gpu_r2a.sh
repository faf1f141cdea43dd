mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r2a_smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q -s > gpurun_out/r2a_pytest.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/r2a_pytest.txt
timeout 300 python bench.py --steps 100 --warmup 5 > gpurun_out/r2a_bench_c2.json 2> gpurun_out/r2a_bench_c2.err
timeout 300 python bench.py --workload c3 --steps 8 --warmup 3 --no-cpu > gpurun_out/r2a_bench_c3.json 2> gpurun_out/r2a_bench_c3.err
timeout 120 python tools/node_trace.py > gpurun_out/r2a_node_trace.txt 2>&1
timeout 120 python tools/tc_trace_c2.py > gpurun_out/r2a_tc_trace_c2.txt 2>&1
tail -5 gpurun_out/r2a_pytest.txt
