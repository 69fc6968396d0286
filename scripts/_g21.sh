set +e
N=$1
mkdir -p gpurun_out/r2m
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 400 $RUN --master-port 29512 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r2m/bench_n$N.json 2> gpurun_out/r2m/bench_n$N.err; echo "rc=$?" >> gpurun_out/r2m/bench_n$N.err
timeout 400 $RUN --master-port 29513 bench.py --gpus $N --steps 10 --warmup 3 --scaling strong > gpurun_out/r2m/bench_n${N}_strong.json 2> gpurun_out/r2m/bench_n${N}_strong.err; echo "rc=$?" >> gpurun_out/r2m/bench_n${N}_strong.err
python - <<PY
import json,glob
for f in sorted(glob.glob('gpurun_out/r2m/bench_n${N}*.json')):
    try:
        d=json.loads([l for l in open(f).read().splitlines() if l.startswith('{')][-1]); print(f, round(d['value']), round(d['ms_per_step'],3), d['scaling'], round(d['e2e']['value']), d['generic_path']['ms_per_step'], d['result_check'])
    except Exception as e: print(f,'ERR',e)
PY
tail -3 gpurun_out/r2m/*.err
