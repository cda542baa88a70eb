mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_active.py -m gpu -x -q 2>&1 | tail -8
python tools/dev_active.py 1024 1250 tensor 2>&1 | grep -E "graph=True|pipelines=3" | tee gpurun_out/active_tensor.log
python tools/dev_active.py 1024 1250 fused 2>&1 | grep -E "graph=True|pipelines=3" | tee gpurun_out/active_fused.log
