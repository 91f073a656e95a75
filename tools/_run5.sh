timeout 600 python -m pytest tests/test_two_pass.py -x -q -m gpu 2>&1 | tail -3
echo "== k=1 lookback pipe"; timeout 200 python tools/kernel_sweep.py --k 1 --only rowids --sels 1e-4,1e-3,1e-2,0.05 2>&1 | grep -v packed | cut -c1-130
echo "== k=1 lookback plain"; CUBIT_LB_NO_PIPE=1 timeout 200 python tools/kernel_sweep.py --k 1 --only rowids --sels 1e-4,1e-2 2>&1 | grep -v packed | cut -c1-130
