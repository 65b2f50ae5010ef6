#!/bin/bash
tag=r02c
run() {
  wl=$1; n=$2; st=$3
  if [ "$n" = 1 ]; then
    timeout 400 python bench.py --workload $wl --gpus 1 --steps $st > gpurun_out/${tag}_${wl}_n1.json 2> gpurun_out/${tag}_${wl}_n1.err
  else
    timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + n)) \
      bench.py --workload $wl --gpus $n --steps $st > gpurun_out/${tag}_${wl}_n$n.json 2> gpurun_out/${tag}_${wl}_n$n.err
  fi
}
run matcha64 8 8
run voc10k 8 2
run tts64 8 20
run matcha64 2 8
