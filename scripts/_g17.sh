set +e
mkdir -p gpurun_out/r2k
timeout 300 ncu --set full --clock-control none -k regex:k_linear_stage3 --launch-skip 80 -c 1 -o gpurun_out/r2k/exp_nk3 -f scripts/_exp_fused_linear.bin 65536 v3 0 > gpurun_out/r2k/ncu_exp.log 2>&1
timeout 300 ncu --set full --clock-control none -k regex:k_linear_stage --launch-skip 55 -c 1 -o gpurun_out/r2k/lib_row2 -f python scripts/_mb_linear.py > gpurun_out/r2k/ncu_lib.log 2>&1
tail -3 gpurun_out/r2k/ncu_exp.log gpurun_out/r2k/ncu_lib.log
ls -la gpurun_out/r2k/*.ncu-rep
