timeout 900 python -m pytest tests/test_wire.py tests/test_host_cpp.py tests/test_duckdb_gpu.py tests/test_round2_gpu.py -x -q -m gpu 2>&1 | tail -4
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 3 --warmup 3 --no-traffic --no-payload24 > gpurun_out/r2_b_2gpu_d.json 2> gpurun_out/r2_b_2gpu_d.err; tail -3 gpurun_out/r2_b_2gpu_d.err
python -c "
import json;d=json.loads(open('gpurun_out/r2_b_2gpu_d.json').read().strip().splitlines()[-1]);print(d['value']/1e12, d['ms_per_step'], d['e2e']['value']/1e12, d['e2e_full_materialize']['value']/1e9, d['e2e_full_materialize_narrow_wire'], d['rowid_gather'])"
