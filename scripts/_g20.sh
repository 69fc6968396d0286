set +e
mkdir -p gpurun_out/r2l
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2l/bench.json 2> gpurun_out/r2l/bench.err; echo "rc=$?" >> gpurun_out/r2l/bench.err
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r2l/bench.json').read().splitlines() if l.startswith('{')][-1])
print('value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'], 'generic', d.get('generic_path'))
r=d['roofline']; print('roofline', r['kernel'][:30], r['achieved'], r['frac'], r['ms_per_attempt'], r['tensor'], r['stage_plus_error_norm'])
print('launches', d['gpu_launches'], d['clocks'], d['result_check'])
PY
tail -3 gpurun_out/r2l/bench.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_linear_stage --launch-skip 600 -c 6 -o gpurun_out/r2l/fused_rows -f python bench.py --steps 1 --warmup 3 --no-device-loop --no-cpu-baseline > gpurun_out/r2l/ncu_fused.log 2>&1
ls -la gpurun_out/r2l/
