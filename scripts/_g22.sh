set +e
mkdir -p gpurun_out/r2n
timeout 2000 python -m pytest tests -m gpu -x -q > gpurun_out/r2n/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2n/pytest_gpu.log
tail -5 gpurun_out/r2n/pytest_gpu.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2n/bench.json 2> gpurun_out/r2n/bench.err; echo "rc=$?" >> gpurun_out/r2n/bench.err
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r2n/bench.json').read().splitlines() if l.startswith('{')][-1])
print('value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'], 'generic', d['generic_path']['ms_per_step'])
r=d['roofline']; print('roofline', r['achieved'], r['frac'], r['ms_per_attempt'], r['stage_plus_error_norm']['frac'], 'cpu', d.get('cpu_baseline',{}).get('value'))
PY
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2n/smoke.log 2>&1; tail -2 gpurun_out/r2n/smoke.log
