timeout 600 python -m pytest tests/test_two_pass.py -x -q -m gpu 2>&1 | tail -3
for pf in 592 0 296 1184; do
echo "== k=1 lookback prefetch $pf"; CUBIT_LB_PREFETCH=$pf timeout 200 python tools/kernel_sweep.py --k 1 --only rowids --sels 1e-4,1e-2,0.05 2>&1 | grep -v packed | cut -c1-130
done
echo "== k=2 lookback"; timeout 200 python tools/kernel_sweep.py --k 2 --only rowids --sels 1e-4,1e-2,0.05 2>&1 | grep -v packed | cut -c1-130
