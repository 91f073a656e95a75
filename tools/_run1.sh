cd $GRAFT_REPO_ROOT
ncu --set full --clock-control none --import-source on -k regex:cubit_scan_kernel -s 2 -c 1 -o gpurun_out/r2_k1_count_m2 python tools/kernel_sweep.py --k 1 --only count --sels 1e-2 --reps 1 > gpurun_out/r2_ncu_k1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:cubit_scan_kernel -s 2 -c 1 -o gpurun_out/r2_k1_rowids_m2 python tools/kernel_sweep.py --k 1 --only rowids --sels 1e-2 --reps 1 >> gpurun_out/r2_ncu_k1.log 2>&1
tail -3 gpurun_out/r2_ncu_k1.log
