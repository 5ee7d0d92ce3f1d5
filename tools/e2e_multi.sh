#!/bin/bash
# N independent processes, one per GPU, each running tools/e2e_trace.py; optional pinning to disjoint core sets.
# usage: tools/e2e_multi.sh N [pin]
N=${1:-8}; PIN=${2:-}
CORES=$(nproc); PER=$((CORES / N))
for i in $(seq 0 $((N - 1))); do
  if [ -n "$PIN" ]; then
    CUDA_VISIBLE_DEVICES=$i taskset -c $((i * PER))-$((i * PER + PER - 1)) python tools/e2e_trace.py 4 > gpurun_out/e2e_multi_$i.log 2>&1 &
  else
    CUDA_VISIBLE_DEVICES=$i python tools/e2e_trace.py 4 > gpurun_out/e2e_multi_$i.log 2>&1 &
  fi
done
wait
for i in $(seq 0 $((N - 1))); do grep "rep 2" gpurun_out/e2e_multi_$i.log; done
