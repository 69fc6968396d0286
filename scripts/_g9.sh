set +e
mkdir -p gpurun_out/r2i
timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_solve.py -m gpu -q --timeout 300 -k "combine or detest_batched or spiral or cubic or fixed or zoo" > gpurun_out/r2i/pytest.log 2>&1; tail -2 gpurun_out/r2i/pytest.log
timeout 600 python scripts/bench_configs.py > gpurun_out/r2i/configs.jsonl 2> gpurun_out/r2i/configs.err
cut -c1-200 gpurun_out/r2i/configs.jsonl
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r2i/dopri8_launches.csv python -c "
import sys; sys.path.insert(0,'scripts'); import bench_configs as b; b.dopri8_roofline()" > gpurun_out/r2i/dopri8.log 2>&1
grep "k_combine_final" gpurun_out/r2i/dopri8_launches.csv | head -3 | cut -c1-60,200-
timeout 120 ./scripts/exp_gemm.bin > gpurun_out/r2i/exp_gemm.txt 2>&1; cat gpurun_out/r2i/exp_gemm.txt
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2i/bench.json 2> gpurun_out/r2i/bench.err; python -c "
import json; d=json.loads(open('gpurun_out/r2i/bench.json').read()); print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['roofline']['combine_plus_error_norm']['frac'])"
