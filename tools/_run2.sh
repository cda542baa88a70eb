mkdir -p gpurun_out
for tm in 0 1 2; do
SPI_B200_WS_TOKEN=$tm SPI_B200_WS_HALVES=2 timeout 300 python tools/dev_halves.py tok$tm 4096 > gpurun_out/tok$tm.log 2>&1
done
SPI_B200_MINB=1 SPI_B200_WS_TOKEN=2 SPI_B200_WS_HALVES=2 timeout 300 python tools/dev_halves.py tok2minb1 4096 > gpurun_out/tok2minb1.log 2>&1
cat gpurun_out/tok*.log
