set +e
mkdir -p gpurun_out/r2h
timeout 1500 python -m pytest tests -m gpu -q --timeout 300 > gpurun_out/r2h/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2h/pytest.log
grep -E "^FAILED|^ERROR|passed|failed" gpurun_out/r2h/pytest.log | tail -15
B="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-device-loop"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_norm --launch-skip 20 -c 1 -o gpurun_out/r2h/k_norm -f $B > gpurun_out/r2h/ncu2.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:k_combine<float, \(int\)5" --launch-skip 20 -c 1 -o gpurun_out/r2h/k_combine5 -f $B > gpurun_out/r2h/ncu3.log 2>&1
for k in k_norm k_combine5; do
  ncu -i gpurun_out/r2h/$k.ncu-rep --page details > gpurun_out/r2h/${k}_details.txt 2>&1
  ncu -i gpurun_out/r2h/$k.ncu-rep --page raw --csv > gpurun_out/r2h/${k}_raw.csv 2>&1
done
rm -f gpurun_out/r2h/*.ncu-rep
timeout 600 python scripts/bench_configs.py > gpurun_out/r2h/configs.jsonl 2> gpurun_out/r2h/configs.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r2h/dopri8_launches.csv python -c "
import sys; sys.path.insert(0,'scripts'); import bench_configs as b; b.dopri8_roofline()" > gpurun_out/r2h/dopri8.log 2>&1
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r2h/bench.json 2> gpurun_out/r2h/bench.err
cat gpurun_out/r2h/configs.jsonl | cut -c1-220
head -c 400 gpurun_out/r2h/bench.json
