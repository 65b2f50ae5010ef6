#!/bin/bash
# Round profiles: (1) per-launch metrics of one whole step, (2) --set full captures of the dominant kernels.
# usage (under gpurun): ./tools/gpu_profile.sh TAG
TAG=${1:-r02}
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,sm__cycles_elapsed.max,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum
timeout 900 ncu --metrics $M --clock-control none --profile-from-start off --csv --log-file gpurun_out/${TAG}_step_metrics.csv python tools/profile_step.py all > /dev/null 2>&1
full() {  # name, step part, kernel regex, skip
  timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:$3 -s $4 -c 1 -f -o gpurun_out/${TAG}_full_$1 python tools/profile_step.py $2 > /dev/null 2>&1
  ncu -i gpurun_out/${TAG}_full_$1.ncu-rep --page raw --csv > gpurun_out/${TAG}_full_$1.raw.csv 2>/dev/null
}
full pair_c32_k11 voc mrf_pair_kernel 16
full pair_c64_k7 voc mrf_pair_kernel 4
full conv_c128_k11 voc conv_bf16_tma_kernel 30
full gemm_split_ffn1 fs2 gemm_split 41
full gemm_split_qkv fs2 gemm_split 44
full attention_dec fs2 relpos_attention 5
ls -la gpurun_out/ | grep $TAG
