#!/bin/bash
# A/B env settings on the vocoder step: ncu launch list per setting (usage: tools/gpu_ab_voc.sh "TAG ENV=.. ENV=.." ...)
for spec in "$@"; do
  set -- $spec; tag=$1; shift
  env "$@" timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/abv_$tag.csv python tools/profile_step.py voc > /dev/null 2>&1
  echo "== $tag: $@"; python tools/launch_summary.py gpurun_out/abv_$tag.csv | grep "conv_bf16\|total"
done
