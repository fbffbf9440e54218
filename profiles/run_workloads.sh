#!/bin/bash
# One bench line per BASELINE.json config, each with its clocks record -> gpurun_out/r02_workloads.jsonl
# (summarised into profiles/r02_workloads.json by profiles/collect_workloads.py).
out=gpurun_out/r02_workloads.jsonl
: > $out
for w in waymo_test kitti_test bev_test waymo_train fpn_waymo mc_uncertainty; do
  python bench.py --workload $w --steps ${STEPS:-20} --warmup 3 >> $out 2>> gpurun_out/r02_workloads.err || echo "{\"workload\": \"$w\", \"failed\": true}" >> $out
done
wc -l $out
