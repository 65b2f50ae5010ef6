#!/bin/bash
# One `gpurun --gpus 8` call: BASELINE config 4 (voc10k, strong scaling), config 5 (matcha64, weak scaling) and config 2
# (tts64, weak scaling) at N = 1, 2, 4, 8, then the stage-4 front-end on 10 000 utterances at 1 and 8 GPUs.
# Outputs: gpurun_out/<tag>_{voc10k,matcha64,tts64}_n{N}.json, <tag>_decode_bench.json
tag=${1:-r02b}
run() {  # workload N steps
  wl=$1; n=$2; st=$3
  if [ "$n" = 1 ]; then
    timeout 600 python bench.py --workload $wl --gpus 1 --steps $st > gpurun_out/${tag}_${wl}_n1.json 2> gpurun_out/${tag}_${wl}_n1.err
  else
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + n)) \
      bench.py --workload $wl --gpus $n --steps $st > gpurun_out/${tag}_${wl}_n$n.json 2> gpurun_out/${tag}_${wl}_n$n.err
  fi
}
for n in 1 2 4 8; do run voc10k $n 2; done
for n in 1 2 4 8; do run matcha64 $n 8; done
for n in 1 8; do run tts64 $n 20; done
timeout 900 python tools/decode_bench.py --utts 10000 --gpus 8 --out gpurun_out/${tag}_decode_bench.json > gpurun_out/${tag}_decode_bench.log 2>&1
