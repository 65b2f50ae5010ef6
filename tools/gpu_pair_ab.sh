#!/bin/bash
# A/B the fused-kernel toggles: each setting in its own process under ncu (per-launch durations)
run() {
  tag=$1; shift
  env "$@" timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:mrf_pair --csv --log-file gpurun_out/ab_$tag.csv python tools/gpu_pair_bench.py > /dev/null 2>&1
  echo "== $tag: $@"; python tools/launch_summary.py gpurun_out/ab_$tag.csv -v | grep "us  " | awk 'NR%2==0{printf "%s ", $2} END{print ""}'
}
for spec in "$@"; do set -- $spec; run "$@"; done
