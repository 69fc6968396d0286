set +e
mkdir -p gpurun_out/r2d
timeout 900 python -m pytest tests -m gpu -q --timeout 300 -k "backprop or plugin" > gpurun_out/r2d/pytest_bp.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2d/pytest_bp.log
grep -E "^FAILED|passed|failed" gpurun_out/r2d/pytest_bp.log | tail -15
timeout 300 ./scripts/exp_combine.bin > gpurun_out/r2d/tma_sweep.txt 2>&1
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r2d/bench.json 2> gpurun_out/r2d/bench.err; echo "bench rc=$?" >> gpurun_out/r2d/bench.err
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-device-loop > gpurun_out/r2d/bench_replay.json 2>> gpurun_out/r2d/bench.err
timeout 600 python scripts/bench_configs.py > gpurun_out/r2d/configs.jsonl 2> gpurun_out/r2d/configs.err
timeout 600 python scripts/timeline.py > gpurun_out/r2d/timeline.txt 2> gpurun_out/r2d/timeline.err
timeout 900 python scripts/detest_sweep.py > gpurun_out/r2d/detest_sweep.jsonl 2> gpurun_out/r2d/detest_sweep.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r2d/launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-device-loop > gpurun_out/r2d/ncu_bench.log 2>&1
tail -3 gpurun_out/r2d/bench.err; head -c 600 gpurun_out/r2d/bench.json; tail -2 gpurun_out/r2d/detest_sweep.jsonl | head -c 600
