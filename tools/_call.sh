mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest.log 2>&1; echo rc=$? >> gpurun_out/pytest.log)
tail -4 gpurun_out/pytest.log
timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -2 gpurun_out/bench.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2>> gpurun_out/bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench.json').read().strip().splitlines()[-1])
print(json.dumps({k:(v if not isinstance(v,dict) else {a:b for a,b in v.items() if a in ('us_per_step','us_per_sweep_at_fixed_point','frac_of_hbm_peak','fp64_TFLOPs','block_updates_per_s')}) for k,v in d['other_configs'].items()}, indent=0))
print(d['value'], d['e2e']['value'], d['quasistatic_events']['value'], d['roofline_stream']['frac'], d['cpu_baseline']['value'], d['clocks'])
print(open('gpurun_out/bench_ref.json').read()[:400])
PY
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r1k_launches.csv python bench.py --realisations 1184 --inner 200 --steps 2 --warmup 1 --no-cpu-baseline --no-other-configs --pipeline 1 > gpurun_out/bench_ncu.log 2>&1
wc -l gpurun_out/r1k_launches.csv
