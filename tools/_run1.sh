set -x
cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_probe_dense.py -x -q > gpurun_out/r2_dense_test.log 2>&1
tail -5 gpurun_out/r2_dense_test.log
timeout 300 python tools/kernel_sweep.py --pack --payload-bits 24 --only fused,agg_only --sels 0.1,0.25,0.5 > gpurun_out/r2_dense_pk24.log 2>&1
timeout 300 python tools/kernel_sweep.py --pack --only fused,agg_only --sels 0.03,0.1,0.25,0.5 > gpurun_out/r2_dense_pk10.log 2>&1
cat gpurun_out/r2_dense_pk24.log gpurun_out/r2_dense_pk10.log
ncu --set full --clock-control none --import-source on -k regex:cubit_probe_dense_kernel -s 1 -c 1 -o gpurun_out/r2_dense_pk10_s05b python tools/kernel_sweep.py --pack --only fused --sels 0.5 --reps 1 > gpurun_out/r2_ncu_dense.log 2>&1
