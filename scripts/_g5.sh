set +e
mkdir -p gpurun_out/r2e
export NCCL_DEBUG=WARN
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 scripts/dist_check.py > gpurun_out/r2e/dist_check_n2.log 2> gpurun_out/r2e/dist_check_n2.err; echo "rc=$?" >> gpurun_out/r2e/dist_check_n2.log
tail -12 gpurun_out/r2e/dist_check_n2.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2e/bench_n2.json 2> gpurun_out/r2e/bench_n2.err; echo "rc=$?" >> gpurun_out/r2e/bench_n2.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 5 --warmup 3 --scaling strong > gpurun_out/r2e/bench_n2_strong.json 2> gpurun_out/r2e/bench_n2_strong.err; echo "rc=$?" >> gpurun_out/r2e/bench_n2_strong.err
tail -3 gpurun_out/r2e/bench_n2.err gpurun_out/r2e/bench_n2_strong.err
python -c "
import json
for f in ('bench_n2','bench_n2_strong'):
    try:
        d=json.loads(open('gpurun_out/r2e/%s.json'%f).read()); print(f, d['value'], d['ms_per_step'], d['scaling'], d['result_check'])
    except Exception as e: print(f, 'ERR', e)
"
timeout 600 python -m pytest tests/test_gpu_dist.py -m gpu -q > gpurun_out/r2e/pytest_dist.log 2>&1; tail -2 gpurun_out/r2e/pytest_dist.log
