#!/bin/bash
# one GPU iteration: pair/op tests, hifigan tests, bench line, vocoder launch list  (usage: tools/gpu_iter.sh TAG [tests])
TAG=$1
TESTS=${2:-"tests/test_mrf_pair_gpu.py tests/test_hifigan_gpu.py"}
timeout 300 python -m pytest $TESTS -x -q 2>&1 | tail -6
timeout 300 python bench.py 2>gpurun_out/${TAG}_bench.err | tail -1 > gpurun_out/${TAG}_bench.json
python -c "
import json; j=json.load(open('gpurun_out/${TAG}_bench.json')); r=j['roofline']
print('audio-s/s', round(j['value']), 'ms/step', round(j['ms_per_step'],2), 'e2e ms', round(j['e2e']['ms_per_step'],2), 'conv ms', round(r['kernel_ms_per_step'],2), 'n', r['launches_per_step'], 'frac', round(r['frac'],3), 'fs2 ms', round(r['fs2_split_gemm']['kernel_ms_per_step'],2), 'launches', j['gpu_launches'])"
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/${TAG}_launches_voc.csv python tools/profile_step.py voc > /dev/null 2>&1
python tools/launch_summary.py gpurun_out/${TAG}_launches_voc.csv -v | grep "ms \|mrf_pair"
