timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_round2_gpu.py -x -q -m gpu -k "delta or Delta or deltas or append or merge or scan_while" 2>&1 | tail -3
timeout 300 python tools/delta_one.py
CUBIT_FORCE_DELTA_KERNEL=1 timeout 200 python tools/kernel_sweep.py --k 10 --only count,rowids --sels 0.1 2>&1 | grep -v packed | cut -c1-130
