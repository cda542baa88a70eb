#!/bin/bash
# dev helper: time the tensor-core actor MLP for several thread-block cluster shapes (one process per shape)
for c in "1,1" "2,1" "1,2" "2,2" "4,1" "1,4" "4,2" "2,4"; do
  echo "== cluster $c"
  SPI_B200_MLP_CLUSTER=$c timeout 120 python tools/dev_mlp_tc.py 11264 2>&1 | grep -E "^tc|Error|error" 
done
