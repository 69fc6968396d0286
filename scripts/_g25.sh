set +e
OUT=gpurun_out/r2p
mkdir -p $OUT
timeout 300 python scripts/_mb_attempt.py > $OUT/mb_attempt_groups.log 2>&1; echo "rc=$?" >> $OUT/mb_attempt_groups.log
cat $OUT/mb_attempt_groups.log | tail -16
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_linear_attempt --launch-skip 31 -c 1 -o $OUT/k_linear_attempt -f python scripts/_mb_attempt.py > $OUT/ncu_attempt.log 2>&1
ncu -i $OUT/k_linear_attempt.ncu-rep --page details > $OUT/k_linear_attempt_details.txt
ncu -i $OUT/k_linear_attempt.ncu-rep --page raw --csv > $OUT/k_linear_attempt_raw.csv
ncu -i $OUT/k_linear_attempt.ncu-rep --page source --csv > $OUT/k_linear_attempt_source.csv 2>/dev/null
ls -la $OUT | tail -8
