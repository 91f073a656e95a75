cd $GRAFT_REPO_ROOT
ncu --set full --clock-control none --import-source on -k "regex:cubit_(scan|probe)" --launch-skip 12 --launch-count 12 -o gpurun_out/r2_step python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-materialize --no-traffic --no-payload24 > gpurun_out/r2_ncu_step.log 2>&1
tail -c 300 gpurun_out/r2_ncu_step.log | tail -2
