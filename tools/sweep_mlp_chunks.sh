#!/bin/bash
# dev helper: actor MLP forward time vs number of independent row-chunk chains
for c in 1 2 4 6 8; do
  echo "== chunks $c"
  SPI_B200_MLP_CHUNKS=$c timeout 120 python tools/dev_mlp_tc.py 11264 2>&1 | grep -E "^tc|Error|error"
done
SPI_B200_MLP_CHUNKS=4 python tools/dev_active.py 1024 1250 2>&1 | tail -2
