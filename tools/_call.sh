mkdir -p gpurun_out
(timeout 300 python -m pytest tests/test_gpu_particles.py tests/test_cpp_host.py -x -q > gpurun_out/pytest_part.log 2>&1; echo rc=$? >> gpurun_out/pytest_part.log)
tail -15 gpurun_out/pytest_part.log
( time timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err ) 2> gpurun_out/bench.time
tail -3 gpurun_out/bench.err; cat gpurun_out/bench.time
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench.json').read().strip().splitlines()[-1])
print(json.dumps(d.get('other_configs'), indent=1))
print(d['value'], d['e2e']['value'], d['quasistatic_events']['value'], d['roofline_stream']['frac'])
PY
