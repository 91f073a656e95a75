set -x
cd $GRAFT_REPO_ROOT
python bench.py --steps 5 --warmup 3 > gpurun_out/r2_b3.json 2> gpurun_out/r2_b3.err
tail -c 600 gpurun_out/r2_b3.err
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2_t4.log 2>&1
tail -5 gpurun_out/r2_t4.log
