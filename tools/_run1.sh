cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_two_pass.py -x -q 2>&1 | tail -3
for k in 1 2 3; do
echo "== k=$k two-pass"; timeout 300 python tools/kernel_sweep.py --k $k --only count,rowids --sels 1e-2,0.5 2>&1 | grep -v packed | cut -c1-140
done
