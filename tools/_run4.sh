timeout 600 python -m pytest tests/test_two_pass.py tests/test_wire.py tests/test_host_cpp.py tests/test_duckdb_gpu.py -x -q -m gpu 2>&1 | tail -5
for k in 1 2; do
echo "== k=$k lookback"; timeout 200 python tools/kernel_sweep.py --k $k --only count,rowids --sels 1e-4,1e-3,1e-2,0.05 2>&1 | grep -v packed | cut -c1-160
echo "== k=$k two-pass"; CUBIT_NO_LOOKBACK=1 timeout 200 python tools/kernel_sweep.py --k $k --only rowids --sels 1e-4,1e-3,1e-2,0.05 2>&1 | grep -v packed | cut -c1-160
done
timeout 200 python tools/drain_sweep.py --threads 8,16 --windows 0,262144,1048576 --sels 0.5,1e-2 > gpurun_out/r2_drain_sweep4.log 2>&1; cut -c1-230 gpurun_out/r2_drain_sweep4.log
timeout 250 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-traffic --no-payload24 > gpurun_out/r2_wire_b4.json 2> gpurun_out/r2_wire_b4.err; tail -3 gpurun_out/r2_wire_b4.err; python -c "
import json;d=json.load(open('gpurun_out/r2_wire_b4.json'));print(d['value']/1e12, d['e2e_full_materialize']['value']/1e9, d['e2e_full_materialize_narrow_wire'])"
