set +e
mkdir -p gpurun_out/r2a
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r2a/smi.txt 2>&1
timeout 300 python scripts/loop_check.py > gpurun_out/r2a/loop_check.log 2>&1; echo "loop_check rc=$?" >> gpurun_out/r2a/loop_check.log
timeout 1500 python -m pytest tests -m gpu -q -x --timeout 300 > gpurun_out/r2a/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2a/pytest.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2a/bench.json 2> gpurun_out/r2a/bench.err; echo "bench rc=$?" >> gpurun_out/r2a/bench.err
timeout 600 python scripts/bench_configs.py > gpurun_out/r2a/configs.jsonl 2> gpurun_out/r2a/configs.err
timeout 600 python scripts/timeline.py > gpurun_out/r2a/timeline.txt 2> gpurun_out/r2a/timeline.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r2a/launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r2a/ncu_bench.log 2>&1
tail -3 gpurun_out/r2a/loop_check.log gpurun_out/r2a/pytest.log gpurun_out/r2a/bench.err
cat gpurun_out/r2a/bench.json | head -c 1500
