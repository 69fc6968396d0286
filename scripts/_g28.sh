set +e
OUT=gpurun_out/r2q
mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest_gpu.log
tail -6 $OUT/pytest_gpu.log
timeout 900 python bench.py --steps 10 --warmup 3 > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"
python - <<'PY'
import json
try:
    d=json.loads([l for l in open('gpurun_out/r2q/bench.json').read().splitlines() if l.startswith('{')][-1])
    print('value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'], 'generic', d.get('generic_path',{}).get('ms_per_step'), 'launches', d['gpu_launches'], 'cpu', d.get('cpu_baseline',{}).get('value'))
    r=d['roofline']; print('roofline', r['bound'], r['achieved'], r['frac'], r['ms_per_attempt'], r.get('replaces'))
except Exception as e:
    print('bench parse failed', e); print(open('gpurun_out/r2q/bench.err').read()[-1500:])
PY
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; tail -2 $OUT/smoke.log
B="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-device-loop --no-fused-controller"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $OUT/launches.csv $B > $OUT/ncu_bench.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_linear_attempt --launch-skip 100 -c 1 -o $OUT/k_linear_attempt -f $B > $OUT/ncu_full.log 2>&1
ncu -i $OUT/k_linear_attempt.ncu-rep --page details > $OUT/k_linear_attempt_details.txt
ncu -i $OUT/k_linear_attempt.ncu-rep --page raw --csv > $OUT/k_linear_attempt_raw.csv
rm -f $OUT/k_linear_attempt.ncu-rep
python scripts/timeline.py > $OUT/timeline.txt 2>&1
ls -la $OUT | tail -12
