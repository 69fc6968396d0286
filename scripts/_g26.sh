set +e
OUT=gpurun_out/r2p
mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_linear.py -m gpu -q -k "attempt" --timeout 300 > $OUT/pytest_attempt.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest_attempt.log
tail -12 $OUT/pytest_attempt.log
timeout 300 python scripts/_mb_attempt.py > $OUT/mb_attempt_groups.log 2>&1; echo "rc=$?" >> $OUT/mb_attempt_groups.log
tail -14 $OUT/mb_attempt_groups.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > $OUT/bench_quick.json 2> $OUT/bench_quick.err; echo "bench rc=$?"
python - <<'PY'
import json
try:
    d=json.loads([l for l in open('gpurun_out/r2p/bench_quick.json').read().splitlines() if l.startswith('{')][-1])
    print('value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'], 'generic', d.get('generic_path',{}).get('ms_per_step'), 'launches', d['gpu_launches'])
    r=d['roofline']; print('roofline', r['bound'], r['achieved'], r['frac'], r['ms_per_attempt'])
except Exception as e:
    print('bench parse failed', e); print(open('gpurun_out/r2p/bench_quick.err').read()[-1500:])
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_linear_attempt --launch-skip 31 -c 1 -o $OUT/k_linear_attempt -f python scripts/_mb_attempt.py > $OUT/ncu_attempt.log 2>&1
ncu -i $OUT/k_linear_attempt.ncu-rep --page details > $OUT/k_linear_attempt_details.txt
ncu -i $OUT/k_linear_attempt.ncu-rep --page raw --csv > $OUT/k_linear_attempt_raw.csv
ncu -i $OUT/k_linear_attempt.ncu-rep --page source --csv > $OUT/k_linear_attempt_source.csv 2>/dev/null
ls -la $OUT | tail -6
