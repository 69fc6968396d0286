set +e
mkdir -p gpurun_out/r2l
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2l/bench_generic.json 2> gpurun_out/r2l/bench_generic.err
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r2l/bench_generic.json').read().splitlines() if l.startswith('{')][-1])
print('generic bench ms', d['ms_per_step'], 'roofline', d['roofline']['frac'], d['roofline']['ms_per_attempt'], d['roofline']['combine_plus_error_norm'])
PY
timeout 900 python -m pytest tests/test_gpu_kernels.py -q -x 2>&1 | tail -3
