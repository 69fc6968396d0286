set +e
OUT=gpurun_out/r2s
mkdir -p $OUT
DIST_CHECK_ADJOINT=0 DIST_CHECK_ROWS_PER_RANK=2048 timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 scripts/dist_check.py > $OUT/dist_check_n2.log 2> $OUT/dist_check_n2.err; echo "dist rc=$?" >> $OUT/dist_check_n2.log
grep -E "^field|DIST_CHECK|rc=" $OUT/dist_check_n2.log | cut -c1-220
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 2 --steps 5 --warmup 3 --no-cpu-baseline > $OUT/bench_n2.json 2> $OUT/bench_n2.err; echo "bench rc=$?"
python - <<'PY'
import json
try:
    d=json.loads([l for l in open('gpurun_out/r2s/bench_n2.json').read().splitlines() if l.startswith('{')][-1])
    print('N=2 value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'], 'check', d['result_check'])
except Exception as e:
    print('bench parse failed', e); print(open('gpurun_out/r2s/bench_n2.err').read()[-2000:])
PY
