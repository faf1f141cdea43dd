# Dev tool: what perturbs the per-step collective at N = 8 (weak-scaled C2 unit): NVML clock polling period, NCCL protocol
mkdir -p gpurun_out
run() { name=$1; shift; env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $((29600 + RANDOM % 200)) bench.py --gpus 8 --steps 300 --warmup 5 --no-extra > gpurun_out/nccl_$name.json 2> gpurun_out/nccl_$name.err; python -c "
import json,sys
try:
    d=json.load(open('gpurun_out/nccl_$name.json')); print('$name', round(d['ms_per_step'],4), round(d['e2e']['ms_per_step'],4))
except Exception as e: print('$name ERR', e)"; }
run poll20ms_a X=1
run nopoll_a TMPNN_BENCH_SAMPLE_PERIOD=0
run poll20ms_b X=1
run nopoll_b TMPNN_BENCH_SAMPLE_PERIOD=0
run poll200ms TMPNN_BENCH_SAMPLE_PERIOD=0.2
run nopoll_ll TMPNN_BENCH_SAMPLE_PERIOD=0 NCCL_PROTO=LL
