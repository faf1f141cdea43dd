#!/bin/bash
# Dev tool: everything the round's committed evidence is made from, in one GPU-box call (outputs under gpurun_out/).
# usage: tools/round_artifacts.sh [tag]
tag=${1:-r01}
out=gpurun_out
mkdir -p $out
set -x
timeout 300 python -m pytest tests -m gpu -q 2>&1 | tail -3 > $out/${tag}_pytest_gpu.txt
timeout 300 python bench.py > $out/${tag}_bench_c2.json 2> $out/${tag}_bench_c2.err
timeout 400 python bench.py --workload c3 --steps 30 > $out/${tag}_bench_c3.json 2> $out/${tag}_bench_c3.err
timeout 300 python bench.py --workload c5 --steps 30 --no-cpu > $out/${tag}_bench_c5.json 2> $out/${tag}_bench_c5.err
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > $out/${tag}_bench_ref.json 2> $out/${tag}_bench_ref.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/${tag}_launches_c2.csv python bench.py --steps 2 --warmup 3 --no-cpu > $out/${tag}_ncu_c2.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/${tag}_launches_c3.csv python bench.py --workload c3 --steps 2 --warmup 3 --no-cpu > $out/${tag}_ncu_c3.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"tc_edge_kernel|node_update" -s 45 -c 4 -f -o $out/${tag}_full_c3 python bench.py --workload c3 --steps 1 --warmup 3 --no-cpu > $out/${tag}_ncu_full_c3.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"tc_edge_kernel|node_update" -s 45 -c 3 -f -o $out/${tag}_full_c2 python bench.py --steps 1 --warmup 3 --no-cpu > $out/${tag}_ncu_full_c2.log 2>&1
ls -la $out | tail -20
