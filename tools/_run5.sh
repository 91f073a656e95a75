timeout 600 python -m pytest tests/test_two_pass.py -x -q -m gpu 2>&1 | tail -3
for k in 1 2; do
echo "== k=$k lookback lane-major"; timeout 200 python tools/kernel_sweep.py --k $k --only rowids --sels 1e-4,1e-3,1e-2,0.05,0.1 2>&1 | grep -v packed | cut -c1-130
done
