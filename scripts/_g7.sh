set +e
N=$1
mkdir -p gpurun_out/r2g
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
if [ "$N" = "8" ]; then
  DIST_CHECK_DUMP_AFTER=200 timeout 400 $RUN --master-port 29511 scripts/dist_check.py > gpurun_out/r2g/dist_check_n$N.log 2> gpurun_out/r2g/dist_check_n$N.err; echo "rc=$?" >> gpurun_out/r2g/dist_check_n$N.log
  DIST_CHECK_ROWS_PER_RANK=65536 DIST_CHECK_ADJOINT=0 DIST_CHECK_DUMP_AFTER=200 timeout 400 $RUN --master-port 29514 scripts/dist_check.py > gpurun_out/r2g/dist_check_big_n$N.log 2> gpurun_out/r2g/dist_check_big_n$N.err; echo "rc=$?" >> gpurun_out/r2g/dist_check_big_n$N.log
  tail -4 gpurun_out/r2g/dist_check_n$N.log gpurun_out/r2g/dist_check_big_n$N.log
  timeout 400 $RUN --master-port 29512 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r2g/bench_n$N.json 2> gpurun_out/r2g/bench_n$N.err; echo "rc=$?" >> gpurun_out/r2g/bench_n$N.err
fi
timeout 400 $RUN --master-port 29513 bench.py --gpus $N --steps 10 --warmup 3 --scaling strong > gpurun_out/r2g/bench_n${N}_strong.json 2> gpurun_out/r2g/bench_n${N}_strong.err; echo "rc=$?" >> gpurun_out/r2g/bench_n${N}_strong.err
timeout 400 $RUN --master-port 29515 bench.py --gpus $N --steps 10 --warmup 3 --scaling strong --no-device-loop > gpurun_out/r2g/bench_n${N}_strong_replay.json 2>> gpurun_out/r2g/bench_n${N}_strong.err
python - <<PY
import json,glob
for f in sorted(glob.glob('gpurun_out/r2g/bench_n${N}*.json')):
    try:
        d=json.loads([l for l in open(f).read().splitlines() if l.startswith('{')][-1]); print(f, round(d['value']), round(d['ms_per_step'],3), d['scaling'], round(d['e2e']['value']), d['result_check'])
    except Exception as e: print(f,'ERR',e)
PY
timeout 300 python -m pytest tests -m gpu -x -q -k "adams" > gpurun_out/r2g/pytest_adams.log 2>&1; tail -2 gpurun_out/r2g/pytest_adams.log
