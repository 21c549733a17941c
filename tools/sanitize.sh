#!/bin/bash
# compute-sanitizer over __graft_entry__.smoke(): one launch (or more) of every hot kernel -- gram_i8_kernel (tcgen05 / TMEM /
# TMA), sweep_tc_kernel (tcgen05 + TMEM + TMA boxes), slgemm_i8_kernel, sweep_pipe_kernel, dense_stream_*, conv_corr9_tma_kernel /
# conv_corr9_strip_kernel (TMA 4-D boxes + mbarrier ring), conv_gram9_tma_kernel (bulk
# copies + mbarrier ring), conv_gram9_nhwc_kernel.  Logs go to gpurun_out/sanitizer_<tool>.log; copy them into profiles/.
set -u
mkdir -p gpurun_out
SAN=/usr/local/cuda/bin/compute-sanitizer
for tool in memcheck racecheck synccheck initcheck; do
  extra=""
  [ "$tool" = "memcheck" ] && extra="--leak-check no"
  timeout 900 $SAN --tool $tool $extra --print-limit 20 --launch-timeout 0 python __graft_entry__.py smoke > gpurun_out/sanitizer_$tool.log 2>&1
  echo "== $tool exit $?"; tail -4 gpurun_out/sanitizer_$tool.log
done
