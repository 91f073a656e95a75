cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_round2_gpu.py -m gpu -x -q -k "delta or random_tables or merge or dml or compressed" > gpurun_out/r2_t7.log 2>&1
tail -4 gpurun_out/r2_t7.log
python tools/delta_one.py > gpurun_out/r2_delta_b.log 2>&1; tail -1 gpurun_out/r2_delta_b.log
