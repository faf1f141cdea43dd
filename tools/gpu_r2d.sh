mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -s > gpurun_out/r2d_pytest.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/r2d_pytest.txt
tail -4 gpurun_out/r2d_pytest.txt
timeout 300 python bench.py --workload c3 --steps 8 --warmup 3 --no-cpu > gpurun_out/r2d_bench_c3.json 2> gpurun_out/r2d_bench_c3.err; echo "c3 rc=$?"
timeout 300 python bench.py --workload c4 --steps 20 --warmup 3 --no-cpu > gpurun_out/r2d_bench_c4.json 2> gpurun_out/r2d_bench_c4.err; echo "c4 rc=$?"
