#!/bin/bash
# usage: tools/run_multi_checks.sh "<world>:<mode>:<launch> ..."   (each check is one torchrun launch of tests/multi_gpu_check.py)
port=29700
for spec in $1; do
  IFS=: read world mode launch <<< "$spec"
  port=$((port+1))
  timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $world --master-addr 127.0.0.1 --master-port $port \
      tests/multi_gpu_check.py $mode $launch 2>&1 | grep -E "MULTI_GPU_CHECK|Error|error" | head -3
  echo "exit $? for $spec"
done
