# Dev tool: the per-step collective's latency at N = 8 under a few NCCL settings (weak-scaled C2 unit)
mkdir -p gpurun_out
run() { name=$1; shift; env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $((29600 + RANDOM % 200)) bench.py --gpus 8 --steps 300 --warmup 5 --no-extra > gpurun_out/nccl_$name.json 2> gpurun_out/nccl_$name.err; python -c "
import json,sys
try:
    d=json.load(open('gpurun_out/nccl_$name.json')); print('$name', round(d['ms_per_step'],4), round(d['e2e']['ms_per_step'],4))
except Exception as e: print('$name ERR', e)"; }
run default X=1
run ll NCCL_PROTO=LL
run ring_ll_2ch NCCL_ALGO=Ring NCCL_PROTO=LL NCCL_MAX_NCHANNELS=2
run nvls_off NCCL_NVLS_ENABLE=0
run simple NCCL_PROTO=Simple
