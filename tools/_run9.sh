timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 8 --steps 10 --warmup 3 --no-traffic --no-payload24 > gpurun_out/r2_b_8gpu_c.json 2> gpurun_out/r2_b_8gpu_c.err; tail -3 gpurun_out/r2_b_8gpu_c.err
python -c "
import json;d=json.loads(open('gpurun_out/r2_b_8gpu_c.json').read().strip().splitlines()[-1]);print(d['value']/1e12, d['ms_per_step'], d['e2e']['value']/1e12, d['e2e_full_materialize']['value']/1e9, d['e2e_full_materialize_narrow_wire'], d['rowid_gather'])"
nproc
