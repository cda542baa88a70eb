mkdir -p gpurun_out
for mb in 2 3 4; do
SPI_B200_MINB=$mb SPI_B200_WS_HALVES=1 timeout 300 python tools/dev_halves.py minb$mb 4096 > gpurun_out/minb$mb.log 2>&1
done
cat gpurun_out/minb*.log
timeout 900 python -m pytest tests/test_readme_bowl.py tests/test_gpu_fullsize.py -m gpu -x -q 2>&1 | tail -5
