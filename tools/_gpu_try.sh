for i in 1 2; do
python bench.py --workload active --steps 6 --warmup 3 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('ftz   ', d['value'], d['ms_per_step'])"
SPI_B200_LIB=tools/_build/libspi_b200_harm2.so python bench.py --workload active --steps 6 --warmup 3 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('no-ftz', d['value'], d['ms_per_step'])"
done
