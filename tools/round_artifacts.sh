#!/bin/bash
# Dev tool: everything the round's committed evidence is made from, in one GPU-box call (outputs under gpurun_out/, which
# must stay below 64 MiB to be copied back: raw ncu reports are reduced to CSV on the box).
# usage: tools/round_artifacts.sh [tag]        (then, in the build container: python tools/summarize_profiles.py <tag> <out-tag>)
tag=${1:-r02}
out=gpurun_out
mkdir -p $out
set -x
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv > $out/${tag}_smi.txt
timeout 600 python -m pytest tests -m gpu -q -s > $out/${tag}_pytest_gpu.txt 2>&1
tail -3 $out/${tag}_pytest_gpu.txt
timeout 600 python bench.py > $out/${tag}_bench_c2.json 2> $out/${tag}_bench_c2.err
timeout 400 python bench.py --workload c3 --steps 30 --no-cpu > $out/${tag}_bench_c3.json 2> $out/${tag}_bench_c3.err
timeout 300 python bench.py --workload c4 --steps 30 --no-cpu > $out/${tag}_bench_c4.json 2> $out/${tag}_bench_c4.err
timeout 300 python bench.py --workload c5 --steps 10 --no-cpu > $out/${tag}_bench_c5.json 2> $out/${tag}_bench_c5.err
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > $out/${tag}_bench_ref.json 2> $out/${tag}_bench_ref.err
export TMPNN_GRAPHS=0      # under the profiler: direct launches (one kernel node = one launch either way)
K='tc_edge_kernel|tc_edge_feat|tc_node_update|tc_head|node_update_mma|head_mma|knn_kernel'
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/${tag}_launches_c2.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-extra > $out/${tag}_ncu_c2.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/${tag}_launches_c3.csv python bench.py --workload c3 --steps 2 --warmup 3 --no-cpu > $out/${tag}_ncu_c3.log 2>&1
for w in c3 c2 c5; do
  extra=""; [ $w = c2 ] && extra="--no-extra"
  timeout 600 ncu --set full --clock-control none -k regex:"$K" -s 36 -c 18 -f -o /tmp/${tag}_full_$w python bench.py --workload $w --steps 1 --warmup 3 --no-cpu $extra > $out/${tag}_ncu_full_$w.log 2>&1
  ncu -i /tmp/${tag}_full_$w.ncu-rep --page raw --csv > $out/${tag}_full_$w.csv 2>/dev/null
done
# one report WITH source correlation, small: the dominant edge kernel on C3
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"tc_edge_kernel<0" -s 3 -c 1 -f -o $out/${tag}_edge_upd_c3 python bench.py --workload c3 --steps 1 --warmup 3 --no-cpu > $out/${tag}_ncu_edge_upd_c3.log 2>&1
du -sh $out; ls -la $out | tail -30
