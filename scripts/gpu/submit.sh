#!/bin/bash
# submit.sh <script> [timeout] [gpus]: run a script under gpurun, retrying while the pod answers busy (exit 3 / transient)
s=$1; t=${2:-900}; n=${3:-1}
for i in $(seq 1 40); do
  out=$(/usr/local/graft/bin/gpurun --timeout $t $( [ "$n" != "1" ] && echo --gpus $n ) -- "bash $s" 2>&1)
  echo "$out" | tail -60
  if echo "$out" | grep -q "status=transient\|status=busy\|no box\|retry in a few minutes"; then sleep 60; continue; fi
  break
done
