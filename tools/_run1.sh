cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_probe_dense.py -x -q 2>&1 | tail -3
timeout 300 python tools/kernel_sweep.py --pack --keep-raw --only fused,agg_only --sels 0.03,0.1,0.25,0.5 2>&1 | grep -v packed_payload | cut -c1-150
for inv in 100 200; do echo inv=$inv; CUBIT_DENSE_MIN_INV=$inv python tools/kernel_sweep.py --pack --keep-raw --only fused,agg_only --sels 0.005,1e-2,0.02 2>&1 | grep -v packed_payload | cut -c1-150; done
echo 24bit; timeout 300 python tools/kernel_sweep.py --pack --payload-bits 24 --only fused --sels 0.1,0.25,0.5 2>&1 | grep -v packed_payload | cut -c1-150
