timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -4 > gpurun_out/r2_t12.log; cat gpurun_out/r2_t12.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 600 python bench.py > gpurun_out/r2_b7.json 2> gpurun_out/r2_b7.err; tail -2 gpurun_out/r2_b7.err
python -c "
import json;d=json.load(open('gpurun_out/r2_b7.json'));print(d['value']/1e12, d['ms_per_step'], d['e2e']['value']/1e12, d['e2e_full_materialize']['value']/1e9, d['e2e_full_materialize_narrow_wire']['value']/1e9, d['roofline']['frac'], d['roofline_probe']['frac'], d['cpu_baseline']['value']/1e9)"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_b7_ref.json 2> gpurun_out/r2_b7_ref.err; tail -c 600 gpurun_out/r2_b7_ref.json
