mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -s > gpurun_out/r2h_pytest.txt 2>&1; tail -3 gpurun_out/r2h_pytest.txt
grep -E "synth600|synth2000|c4 batch|synth300:" gpurun_out/r2h_pytest.txt
timeout 300 python bench.py --workload c5 --steps 5 --warmup 3 --no-cpu > gpurun_out/r2h_bench_c5.json 2> gpurun_out/r2h_bench_c5.err; echo "c5 rc=$?"
timeout 300 python bench.py --workload c3 --steps 10 --warmup 3 --no-cpu > gpurun_out/r2h_bench_c3.json 2> gpurun_out/r2h_bench_c3.err; echo "c3 rc=$?"
timeout 300 python bench.py --steps 200 --warmup 5 --no-cpu --no-extra > gpurun_out/r2h_bench_c2.json 2> gpurun_out/r2h_bench_c2.err; echo "c2 rc=$?"
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 3 python -m pytest tests -m gpu -x -q -k "c4_shaped or long_proteins or ragged_batch or log_probs or cfg_variants or graph_replay" > gpurun_out/r2h_memcheck.txt 2>&1; echo "memcheck rc=$?"; tail -5 gpurun_out/r2h_memcheck.txt
python -c "
import sys; sys.path.insert(0,'.')
import __graft_entry__ as g; g.smoke()" > gpurun_out/r2h_smoke.txt 2>&1; tail -3 gpurun_out/r2h_smoke.txt
