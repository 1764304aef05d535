#!/bin/bash
# usage: tools/gpurun_retry.sh <timeout_s> <gpus> '<command>'  -- retries while the pod answers busy / transient (nothing is charged then)
T=$1; G=$2; shift 2
for attempt in $(seq 1 40); do
  if [ "$G" = "1" ]; then /usr/local/graft/bin/gpurun --timeout "$T" -- "$@"; else /usr/local/graft/bin/gpurun --gpus "$G" --timeout "$T" -- "$@"; fi
  rc=$?
  st=$(python -c "import json;print(json.load(open('/root/repo/gpurun_out/.last_call.json')).get('status'))" 2>/dev/null)
  if [ "$st" != "transient" ] && [ "$rc" != "3" ]; then exit $rc; fi
  echo "[retry] attempt $attempt: $st rc=$rc; sleeping 120 s"
  sleep 120
done
exit 3
