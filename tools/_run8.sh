# ncu --set full of the look-back scan (k = 1, s = 1e-3) and of the two wire kernels (one drain window)
timeout 200 python tools/kernel_sweep.py --k 1 --only rowids --sels 1e-3 --reps 1 > /dev/null 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:lookback -c 2 -f -o gpurun_out/r2_lookback python tools/kernel_sweep.py --k 1 --only rowids --sels 1e-3 --reps 1 > gpurun_out/ncu_lb.log 2>&1; tail -2 gpurun_out/ncu_lb.log | cut -c1-200
timeout 600 ncu --set full --clock-control none --import-source on -k regex:cubit_wire -c 4 -f -o gpurun_out/r2_wire python tools/drain_sweep.py --rows 100000000 --threads 1 --windows 262144 --sels 0.5 --reps 1 > gpurun_out/ncu_wire.log 2>&1; tail -2 gpurun_out/ncu_wire.log | cut -c1-200
ls -la gpurun_out/*.ncu-rep | tail -3
