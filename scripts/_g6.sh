set +e
mkdir -p gpurun_out/r2f
B="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-device-loop"
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:k_combine_final<float" --launch-skip 20 -c 1 -o gpurun_out/r2f/k_combine_final -f $B > gpurun_out/r2f/ncu1.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:k_norm<float, 0" --launch-skip 20 -c 1 -o gpurun_out/r2f/k_norm -f $B > gpurun_out/r2f/ncu2.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:k_combine<float, 5" --launch-skip 20 -c 1 -o gpurun_out/r2f/k_combine5 -f $B > gpurun_out/r2f/ncu3.log 2>&1
for k in k_combine_final k_norm k_combine5; do
  ncu -i gpurun_out/r2f/$k.ncu-rep --page details > gpurun_out/r2f/${k}_details.txt 2>&1
  ncu -i gpurun_out/r2f/$k.ncu-rep --page raw --csv > gpurun_out/r2f/${k}_raw.csv 2>&1
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r2f/dopri8_launches.csv python -c "
import sys; sys.path.insert(0,'scripts'); import bench_configs as b; b.dopri8_roofline()" > gpurun_out/r2f/dopri8.log 2>&1
ls -la gpurun_out/r2f | head -20
timeout 600 python scripts/timeline.py > gpurun_out/r2f/timeline.txt 2> gpurun_out/r2f/timeline.err
timeout 600 python -m pytest tests -m gpu -q --timeout 300 -k "backprop or many_parameter" > gpurun_out/r2f/pytest_sub.log 2>&1; tail -3 gpurun_out/r2f/pytest_sub.log
