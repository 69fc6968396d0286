set +e
mkdir -p gpurun_out/r2j
timeout 1500 python -m pytest tests -m gpu -q --timeout 300 > gpurun_out/r2j/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2j/pytest.log
grep -E "^FAILED|^ERROR|passed|failed" gpurun_out/r2j/pytest.log | tail -15
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r2j/bench.json 2> gpurun_out/r2j/bench.err
python -c "
import json; d=json.loads(open('gpurun_out/r2j/bench.json').read()); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d['roofline']['combine_plus_error_norm']['frac'], d['cpu_baseline'])"
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2j/bench_reference.json 2>> gpurun_out/r2j/bench.err; head -c 300 gpurun_out/r2j/bench_reference.json
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2j/smoke.log 2>&1; tail -1 gpurun_out/r2j/smoke.log
