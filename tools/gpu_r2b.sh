mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -s > gpurun_out/r2b_pytest.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/r2b_pytest.txt
timeout 300 python bench.py --steps 200 --warmup 5 --no-cpu > gpurun_out/r2b_bench_c2.json 2> gpurun_out/r2b_bench_c2.err
TMPNN_LIB=dev TMPNN_SMALL_SETS=0 timeout 300 python bench.py --steps 200 --warmup 5 --no-cpu > gpurun_out/r2b_bench_c2_sets2.json 2> gpurun_out/r2b_bench_c2_sets2.err
timeout 120 python tools/node_trace.py > gpurun_out/r2b_node_trace.txt 2>&1
timeout 120 python tools/tc_trace_c2.py > gpurun_out/r2b_tc_trace_c2.txt 2>&1
tail -3 gpurun_out/r2b_pytest.txt
