# Dev tool: the 1/2/4/8-GPU scaling curve on one 8-GPU box (outputs gpurun_out/<tag>_bench_n*.json)
tag=${1:-r02}
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/${tag}_gpus.txt
for n in 8 4 2; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29520 + n)) bench.py --gpus $n --steps 100 --warmup 5 > gpurun_out/${tag}_bench_n$n.json 2> gpurun_out/${tag}_bench_n$n.err; echo "n$n rc=$?"
done
timeout 300 python bench.py --gpus 1 --steps 100 --warmup 5 --no-cpu --no-extra > gpurun_out/${tag}_bench_n1.json 2> gpurun_out/${tag}_bench_n1.err; echo "n1 rc=$?"
timeout 300 python bench.py --workload c5 --steps 5 --warmup 3 --no-cpu > gpurun_out/${tag}_bench_c5_n1.json 2> gpurun_out/${tag}_bench_c5_n1.err; echo "c5 n1 rc=$?"
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/*_bench_n*.json")):
    try: d = json.load(open(f))
    except Exception as e: print(f, "ERR", e); continue
    print(f, d["n_gpus"], round(d["ms_per_step"], 4), "%.3e" % d["value"], "e2e", round(d["e2e"]["ms_per_step"], 4), d.get("sharded_c5", {}).get("ms_per_step"))
PY
