mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -s > gpurun_out/r2f_pytest.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/r2f_pytest.txt
tail -3 gpurun_out/r2f_pytest.txt; grep -E "synth600|synth2000|c4 batch" gpurun_out/r2f_pytest.txt
timeout 200 python tools/tc_node_trace.py 50 > gpurun_out/r2f_tc_node_trace.txt 2>&1; tail -2 gpurun_out/r2f_tc_node_trace.txt
timeout 300 python bench.py --workload c3 --steps 8 --warmup 3 --no-cpu > gpurun_out/r2f_bench_c3.json 2> gpurun_out/r2f_bench_c3.err; echo "c3 rc=$?"
timeout 300 python bench.py --workload c4 --steps 20 --warmup 3 --no-cpu > gpurun_out/r2f_bench_c4.json 2> gpurun_out/r2f_bench_c4.err; echo "c4 rc=$?"
