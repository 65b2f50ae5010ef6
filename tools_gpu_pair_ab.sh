#!/bin/bash
# A/B the fused-kernel toggles: each setting in its own process under ncu (per-launch durations)
run() {
  tag=$1; shift
  env "$@" timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:mrf_pair --csv --log-file gpurun_out/ab_$tag.csv python tools_gpu_pair_bench.py > /dev/null 2>&1
  echo "== $tag: $@"; python tools_launch_summary.py gpurun_out/ab_$tag.csv -v | grep "us  " | awk 'NR%2==0{printf "%s %s | ", $2, $4} END{print ""}'
}
run base X=1
run lag0 JATTS_B200_PAIR_LAG=0
run lag1 JATTS_B200_PAIR_LAG=1
run poll0 JATTS_B200_PAIR_POLL=0
run sa4 JATTS_B200_PAIR_SA=4
run stream4 JATTS_B200_PAIR_STREAM=4
