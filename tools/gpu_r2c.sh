mkdir -p gpurun_out
timeout 600 python bench.py --steps 100 --warmup 5 > gpurun_out/r2c_bench_c2.json 2> gpurun_out/r2c_bench_c2.err; echo "c2 rc=$?"
timeout 300 python bench.py --workload c4 --steps 20 --warmup 3 --no-cpu > gpurun_out/r2c_bench_c4.json 2> gpurun_out/r2c_bench_c4.err; echo "c4 rc=$?"
timeout 300 python bench.py --workload c5 --steps 5 --warmup 3 --no-cpu > gpurun_out/r2c_bench_c5.json 2> gpurun_out/r2c_bench_c5.err; echo "c5 rc=$?"
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r2c_bench_ref.json 2> gpurun_out/r2c_bench_ref.err; echo "ref rc=$?"
tail -3 gpurun_out/r2c_bench_c2.err
