#!/bin/bash
# needs a library built with -DSPI_B200_MLP_DEV (NVCC_FLAGS in spi_active_b200/_lib.py): isolation runs of the actor kernel
for l in 7 1 2 4; do for d in 0 1 2 3; do
  echo "== layers $l dbg $d"
  SPI_B200_MLP_CLUSTER=1,1 SPI_B200_MLP_LAYERS=$l SPI_B200_MLP_DBG=$d timeout 120 python tools/dev_mlp_tc.py 11264 2>&1 | grep -E "^tc|Error|error"
done; done
