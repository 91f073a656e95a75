timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -4 > gpurun_out/r2_t11.log; cat gpurun_out/r2_t11.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 600 python bench.py > gpurun_out/r2_b6.json 2> gpurun_out/r2_b6.err; tail -2 gpurun_out/r2_b6.err
python -c "
import json;d=json.load(open('gpurun_out/r2_b6.json'));print(d['value']/1e12, d['ms_per_step'], d['e2e']['value']/1e12, d['e2e_full_materialize']['value']/1e9, d['e2e_full_materialize_narrow_wire']['value']/1e9, d['roofline']['frac'], d['roofline_probe']['frac'], d['cpu_baseline']['value']/1e9)"
timeout 300 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-materialize --no-traffic --no-payload24 > /dev/null 2>&1 && timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_c.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-materialize --no-traffic --no-payload24 > gpurun_out/ncu_c.log 2>&1; tail -2 gpurun_out/ncu_c.log | cut -c1-200
