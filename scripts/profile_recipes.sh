#!/bin/bash
# The commands behind profiles/r2_* (run on a B200 box from the repo root, e.g. `gpurun -- 'bash scripts/profile_recipes.sh'`).
# Numbers printed under ncu are never bench values.
set +e
OUT=${OUT:-gpurun_out/profile}
mkdir -p $OUT
# kernels inside the device-side while loop are invisible to ncu's kernel-level profiling: profile the host-replay mode
B="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-device-loop"
# 1. launch list (per-kernel durations of a warm solve)
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file $OUT/launches.csv $B > $OUT/ncu_bench.log 2>&1
# 2. full captures of the three dominant kernels
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:k_combine_final<float" --launch-skip 20 -c 1 -o $OUT/k_combine_final -f $B
ncu --set full --clock-control none --import-source on -k regex:k_norm --launch-skip 20 -c 1 -o $OUT/k_norm -f $B
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:k_combine<float, \(int\)5" --launch-skip 20 -c 1 -o $OUT/k_combine5 -f $B
for k in k_combine_final k_norm k_combine5; do
  ncu -i $OUT/$k.ncu-rep --page details > $OUT/${k}_details.txt
  ncu -i $OUT/$k.ncu-rep --page raw --csv > $OUT/${k}_raw.csv
done
# 3. the other configs, the small-state timelines, the DETEST sweep, the TMA experiment
python scripts/bench_configs.py > $OUT/configs.jsonl
python scripts/timeline.py > $OUT/timeline.txt
python scripts/detest_sweep.py > $OUT/detest_sweep.jsonl
nvcc -gencode arch=compute_100a,code=sm_100a -O3 --fmad=false -o $OUT/exp_combine scripts/exp_combine.cu && $OUT/exp_combine > $OUT/tma_sweep.txt
# 4. multi-GPU (N = 2, 4, 8):  torchrun --nproc-per-node N scripts/dist_check.py ;  bench.py --gpus N [--scaling strong]

# 5. the whole-attempt kernel (csrc/tdq_attempt.cu; profiles/r2_*attempt*).  Under ncu the controller step must be a launch of its
#    own (--no-fused-controller): a kernel that carries a device-runtime call is skipped by kernel-level profiling.
#   python bench.py --steps 10 --warmup 3                                  > r2_bench_n1_attempt.json
#   python scripts/_mb_attempt.py                                          > r2_mb_attempt_variants.txt   (stand-alone launches, bitwise check first)
#   B2="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-device-loop --no-fused-controller"
#   ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file r2_launches_bench_attempt.csv $B2
#   ncu --set full --clock-control none --import-source on -k regex:k_linear_attempt --launch-skip 100 -c 1 -o k_linear_attempt -f $B2
#   torchrun --nproc-per-node 2 bench.py --gpus 2 --steps 5 --warmup 3 --no-cpu-baseline                      > r2_bench_n2_attempt.json
#   DIST_CHECK_ADJOINT=0 DIST_CHECK_ROWS_PER_RANK=2048 torchrun --nproc-per-node 2 scripts/dist_check.py      > r2_dist_check_n2_attempt.log
