mkdir -p gpurun_out
timeout 800 compute-sanitizer --tool racecheck --print-limit 5 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "eval_candidates_matches_oracle or rollout_states or motor_models or mask_and_ragged" > gpurun_out/racecheck_parity_r2c.log 2>&1; echo racecheck rc=$?
tail -5 gpurun_out/racecheck_parity_r2c.log
timeout 600 compute-sanitizer --tool memcheck --print-limit 5 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/memcheck_parity_r2c.log 2>&1; echo memcheck rc=$?
tail -5 gpurun_out/memcheck_parity_r2c.log
